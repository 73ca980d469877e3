"""Stand-in for the ``mpi4py`` package (TEST INFRASTRUCTURE, see MPI.py)."""
__version__ = '0.0-fake'
