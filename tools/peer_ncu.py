"""One process, two GPUs: the fused FFT + redistribution kernel of the 2-GPU 1024^3 step (stage 1: transform
along axis 1 of the (S/2, S, S) block, last pass storing each point into the owner's (S, S/2, S) window) with
rank 0 played on cuda:0 and rank 1's window living on cuda:1 -- real NVLink traffic from a single process, so
that `ncu` can wrap it (it cannot wrap a torchrun job).  Prints the event time and the remote GB/s.

    python tools/peer_ncu.py [--size 1024] [--reps 5]
    ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum \\
        -k regex:peer -c 2 python tools/peer_ncu.py --reps 1
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    import torch
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200._lib import TransferHandle, Plan
    if torch.cuda.device_count() < 2:
        print("needs two GPUs")
        return
    S, p = args.size, 2
    # make device 1's memory reachable from kernels on device 0 (torch enables peer access on the first copy)
    assert torch.cuda.can_device_access_peer(0, 1)
    x1 = torch.zeros(1 << 20, device='cuda:1')
    x0 = torch.zeros(1 << 20, device='cuda:0')
    x0.copy_(x1)
    x1.copy_(x0)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    torch.cuda.set_device(0)
    shape = (S, S, S)
    src_shape = (S // p, S, S)          # rank 0's block: split along axis 0, full along axis 1 (transformed)
    dst_shape = (S, S // p, S)          # after the transfer: full along axis 0, split along axis 1
    src = torch.view_as_complex(torch.rand(src_shape + (2,), dtype=torch.float64, device='cuda:0'))
    dst0 = torch.zeros(dst_shape, dtype=torch.complex128, device='cuda:0')
    dst1 = torch.zeros(dst_shape, dtype=torch.complex128, device='cuda:1')

    class Rank0(object):
        ranks = (0, 1)

        def Get_size(self):
            return p

        def Get_rank(self):
            return 0
    h = TransferHandle(Rank0(), shape, 16, src_shape, 1, dst_shape, 0, exchange=False)
    plan = Plan(src_shape, src_shape, (1,), [-1], 8)
    assert plan.can_scatter(h, 0)
    ptrs = [dst0.data_ptr(), dst1.data_ptr()]
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        plan.execute_scatter(src.data_ptr(), 0, 1.0, h, 0, ptrs, sync=False)
    run()
    torch.cuda.synchronize(0)
    e0.record(stream)
    for _ in range(args.reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    ms = e0.elapsed_time(e1) / args.reps
    remote = src.numel() * 16 * (p - 1) / p
    print("fused stage, %d^3 complex128 on 2 GPUs: %.3f ms per launch, %.2f GB remote -> %.1f GB/s over NVLink, "
          "local block %.2f GB in + %.2f GB out" % (S, ms, remote / 1e9, remote / ms / 1e6, src.numel() * 16 / 1e9,
                                                     src.numel() * 16 / 1e9))
    # parity of what landed on the peer: rows [0, S/2) of axis 0 of rank 1's window = FFT along axis 1, second half of it
    chk = torch.fft.fft(src[:2], dim=1)[:, S // 2:, :]
    got = dst1[:2].to('cuda:0')
    err = float((got - chk).abs().max() / chk.abs().max())
    print("max rel err of the block that crossed NVLink: %.3g" % err)
    assert err < 1e-12
    plan.destroy()
    h.destroy()


if __name__ == '__main__':
    main()
