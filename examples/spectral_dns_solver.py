"""Taylor-Green vortex in a triply periodic box, pseudo-spectral Navier-Stokes with
RK4 -- the end-to-end known-answer test of the reference
(/root/reference/examples/spectral_dns_solver.py; energy after 10 steps at 64^3 =
0.124953117517, asserted to 7 decimals at :129), run on the device arrays of this
package.  Same structure and names as the reference script; what differs is that
every array lives in HBM (numpy ufuncs become the elementwise operators of
DeviceArray / a torch reduction) and that the 3/2-rule padded variant, which the
reference leaves commented out, is a switch:

    python examples/spectral_dns_solver.py [--padding] [--M 6]
    torchrun --nproc-per-node 4 examples/spectral_dns_solver.py
"""
import argparse
import os
import sys
from time import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def solve(M=6, padding=False, T=0.1, dt=0.01, nu=0.000625, comm=None):
    """Returns the kinetic energy at time T (rank 0; None elsewhere)."""
    import torch
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200 import PFFT, newDistArray, DeviceArray
    comm = B.init() if comm is None else comm

    N = [2 ** M, 2 ** M, 2 ** M]
    L = np.array([2 * np.pi, 4 * np.pi, 4 * np.pi], dtype=float)

    FFT = PFFT(comm, N, collapse=False)
    FFT_pad = PFFT(comm, N, padding=[1.5, 1.5, 1.5]) if padding else FFT

    U = newDistArray(FFT, False, rank=1)
    U_hat = newDistArray(FFT, rank=1)
    P_hat = newDistArray(FFT)
    U_hat0 = newDistArray(FFT, rank=1)
    U_hat1 = newDistArray(FFT, rank=1)
    a = [1. / 6., 1. / 3., 1. / 3., 1. / 6.]
    b = [0.5, 0.5, 1.]
    dU = newDistArray(FFT, rank=1)
    U_pad = newDistArray(FFT_pad, False, rank=1)
    curl_pad = newDistArray(FFT_pad, False, rank=1)

    def get_local_mesh(FFT, L):
        X = np.ogrid[FFT.local_slice(False)]
        Ng = FFT.global_shape()
        return [np.broadcast_to(x * L[i] / Ng[i], FFT.shape(False)) for i, x in enumerate(X)]

    def get_local_wavenumbermesh(FFT, L):
        s = FFT.local_slice()
        Ng = FFT.global_shape()
        k = [np.fft.fftfreq(n, 1. / n).astype(int) for n in Ng[:-1]]
        k.append(np.fft.rfftfreq(Ng[-1], 1. / Ng[-1]).astype(int))
        K = [ki[si] for ki, si in zip(k, s)]
        Ks = np.meshgrid(*K, indexing='ij', sparse=True)
        Lp = 2 * np.pi / L
        return [np.broadcast_to(k * Lp[i], FFT.shape(True)) for i, k in enumerate(Ks)]

    X = get_local_mesh(FFT, L)
    Kh = np.array(get_local_wavenumbermesh(FFT, L)).astype(float)
    K2h = np.sum(Kh * Kh, 0, dtype=float)
    dev = lambda h: DeviceArray(torch.from_numpy(np.ascontiguousarray(h)).cuda())
    K, K2 = dev(Kh), dev(K2h)
    K_over_K2 = dev(Kh / np.where(K2h == 0, 1, K2h))

    def cross(x, y, z):
        FFT_pad.forward(x[1] * y[2] - x[2] * y[1], z[0])
        FFT_pad.forward(x[2] * y[0] - x[0] * y[2], z[1])
        FFT_pad.forward(x[0] * y[1] - x[1] * y[0], z[2])
        return z

    def compute_curl(x, z):
        FFT_pad.backward(1j * (K[0] * x[1] - K[1] * x[0]), z[2])
        FFT_pad.backward(1j * (K[2] * x[0] - K[0] * x[2]), z[1])
        FFT_pad.backward(1j * (K[1] * x[2] - K[2] * x[1]), z[0])
        return z

    def compute_rhs(rhs):
        for j in range(3):
            FFT_pad.backward(U_hat[j], U_pad[j])
        compute_curl(U_hat, curl_pad)
        rhs = cross(U_pad, curl_pad, rhs)
        P_hat[...] = (rhs * K_over_K2).tensor.sum(0)
        rhs -= P_hat * K
        rhs -= nu * K2 * U_hat
        return rhs

    U[0] = np.sin(X[0]) * np.cos(X[1]) * np.cos(X[2])
    U[1] = -np.cos(X[0]) * np.sin(X[1]) * np.cos(X[2])
    U[2] = 0
    for i in range(3):
        FFT.forward(U[i], U_hat[i])

    t, t0 = 0.0, time()
    while t < T - 1e-8:
        t += dt
        U_hat1[...] = U_hat
        U_hat0[...] = U_hat
        for rk in range(4):
            dU = compute_rhs(dU)
            if rk < 3:
                U_hat[...] = U_hat0 + b[rk] * dt * dU
            U_hat1 += a[rk] * dt * dU
        U_hat[...] = U_hat1
        for i in range(3):
            FFT.backward(U_hat[i], U[i])

    k = comm.reduce((U * U).sum() / N[0] / N[1] / N[2] / 2)
    torch.cuda.synchronize()
    elapsed = time() - t0
    FFT.destroy()
    if FFT_pad is not FFT:
        FFT_pad.destroy()
    if comm.Get_rank() == 0:
        print('Time = {}'.format(elapsed))
        return float(k)
    return None


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--M', type=int, default=6)
    ap.add_argument('--padding', action='store_true')
    args = ap.parse_args()
    k = solve(args.M, args.padding)
    if k is not None:
        print('Energy = {:.12f}'.format(k))
        if args.M == 6:
            assert round(k - 0.124953117517, 7) == 0, k
