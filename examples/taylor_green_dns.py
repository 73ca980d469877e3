"""Pseudo-spectral Navier-Stokes on device arrays: the Taylor-Green vortex in a
periodic box, classical RK4 in Fourier space, rotational form of the non-linear
term.  This is the end-to-end known-answer test of the hot path: the reference
ships the same experiment (/root/reference/examples/spectral_dns_solver.py) and
asserts the kinetic energy after ten steps at 64^3, 0.124953117517, to seven
decimals (:129).  Everything between the initial condition and the final energy
lives in HBM: PFFT r2c/c2r + c2c stages, rank-1 DistArrays, elementwise torch
arithmetic on the arrays' tensors.

    python examples/taylor_green_dns.py [--log2n 6] [--dealias]
    torchrun --nproc-per-node 4 examples/taylor_green_dns.py

--dealias runs the products on a 3/2-padded grid (PFFT(padding=[1.5]*3)).
"""
import argparse
import os
import sys
from time import perf_counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KNOWN_ENERGY_64 = 0.124953117517      # reference examples/spectral_dns_solver.py:129
BOX = (2 * np.pi, 4 * np.pi, 4 * np.pi)


class TaylorGreen(object):
    """State and right-hand side of du^/dt = P[ (u x w)^ ] - nu k^2 u^  (P = projection
    onto divergence-free fields) for one rank's block of the spectrum."""

    def __init__(self, comm, log2n=6, dealias=False, nu=0.000625):
        import torch
        import mpi4py_fft_b200 as B
        self.torch, self.B, self.comm, self.nu = torch, B, comm, nu
        n = 2 ** log2n
        self.n = (n, n, n)
        self.fft = B.PFFT(comm, self.n, collapse=False)
        self.fft_nl = B.PFFT(comm, self.n, padding=[1.5] * 3) if dealias else self.fft
        new = B.newDistArray
        self.u = new(self.fft, False, rank=1)                 # velocity, physical space
        self.u_hat = new(self.fft, True, rank=1)              # velocity, spectral space
        self.rhs = new(self.fft, True, rank=1)
        self.u_nl = new(self.fft_nl, False, rank=1)           # velocity / vorticity on the product grid
        self.w_nl = new(self.fft_nl, False, rank=1)
        self.k, self.k2, self.k_over_k2 = self._wavenumbers()

    def _wavenumbers(self):
        """this rank's wavenumber vectors as device tensors (3, n0', n1', n2')"""
        torch = self.torch
        sl = self.fft.local_slice(True)
        full = [np.fft.fftfreq(m, 1.0 / m) for m in self.n[:-1]] + [np.fft.rfftfreq(self.n[-1], 1.0 / self.n[-1])]
        local = [f[s] * (2 * np.pi / BOX[i]) for i, (f, s) in enumerate(zip(full, sl))]
        grid = np.stack(np.meshgrid(*local, indexing='ij'))
        k2 = (grid ** 2).sum(0)
        safe = np.where(k2 == 0, 1.0, k2)
        put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return put(grid), put(k2), put(grid / safe)

    def initial_condition(self):
        sl = self.fft.local_slice(False)
        x, y, z = np.meshgrid(*[np.arange(s.start, s.stop) * BOX[i] / self.n[i] for i, s in enumerate(sl)],
                              indexing='ij')
        self.u[0] = np.sin(x) * np.cos(y) * np.cos(z)
        self.u[1] = -np.cos(x) * np.sin(y) * np.cos(z)
        self.u[2] = 0
        for c in range(3):
            self.fft.forward(self.u[c], self.u_hat[c])

    def right_hand_side(self, u_hat, out):
        """out <- rhs(u_hat); both are rank-1 spectral DistArrays"""
        B, k = self.B, self.k
        uh = u_hat.tensor
        # vorticity in spectral space, then both fields to the (possibly padded) physical grid
        curl = (k[1] * uh[2] - k[2] * uh[1], k[2] * uh[0] - k[0] * uh[2], k[0] * uh[1] - k[1] * uh[0])
        for c in range(3):
            self.fft_nl.backward(u_hat[c], self.u_nl[c])
            self.fft_nl.backward(B.DeviceArray(1j * curl[c]), self.w_nl[c])
        u, w = self.u_nl.tensor, self.w_nl.tensor
        cross = (u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0])
        for c in range(3):
            self.fft_nl.forward(B.DeviceArray(cross[c]), out[c])
        o = out.tensor
        pressure = (o * self.k_over_k2).sum(0)
        o -= pressure * k
        o -= self.nu * self.k2 * uh
        return out

    def advance(self, dt):
        """one classical Runge-Kutta step in spectral space"""
        weights, nodes = (1 / 6, 1 / 3, 1 / 3, 1 / 6), (0.5, 0.5, 1.0)
        start = self.u_hat.tensor.clone()
        total = start.clone()
        for stage in range(4):
            self.right_hand_side(self.u_hat, self.rhs)
            if stage < 3:
                self.u_hat.tensor.copy_(start + nodes[stage] * dt * self.rhs.tensor)
            total += weights[stage] * dt * self.rhs.tensor
        self.u_hat.tensor.copy_(total)

    def energy(self):
        for c in range(3):
            self.fft.backward(self.u_hat[c], self.u[c])
        local = float((self.u.tensor ** 2).sum().item()) / float(np.prod(self.n)) / 2
        return self.comm.allreduce(local)

    def close(self):
        self.fft.destroy()
        if self.fft_nl is not self.fft:
            self.fft_nl.destroy()


def solve(log2n=6, dealias=False, t_end=0.1, dt=0.01, comm=None):
    """kinetic energy at t_end (the same number on every rank)"""
    import torch
    import mpi4py_fft_b200 as B
    comm = B.init() if comm is None else comm
    flow = TaylorGreen(comm, log2n, dealias)
    flow.initial_condition()
    t0 = perf_counter()
    for _ in range(int(round(t_end / dt))):
        flow.advance(dt)
    e = flow.energy()
    torch.cuda.synchronize()
    if comm.Get_rank() == 0:
        print("%d^3, %d steps: %.3f s, energy %.12f" % (2 ** log2n, int(round(t_end / dt)), perf_counter() - t0, e))
    flow.close()
    return e


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--log2n', type=int, default=6)
    ap.add_argument('--dealias', action='store_true')
    args = ap.parse_args()
    energy = solve(args.log2n, args.dealias)
    if args.log2n == 6:
        assert round(energy - KNOWN_ENERGY_64, 7) == 0, energy
