"""Stage the UNMODIFIED reference package under oracle/_ref/ (git-ignored) so it
can be imported next to the fake ``mpi4py`` of oracle/fakempi -- the Python
analogue of `pip install --target`, which cannot work here because the
reference's setup.py refuses to build without libfftw3 (setup.py:78-80).

Only ``mpi4py_fft/fftw/utilities.pyx`` needs compiling (numpy headers only); the
FFTW-backed ``fftw_xfftn`` extensions cannot be built (no libfftw3 in the
image), so the staged reference runs with its own ``numpy``/``scipy`` serial
backends (libfft.py:81-102,128-144).  Nothing is copied into tracked paths.
Run in the build container only: /root/reference does not exist on the GPU box,
the staged copy travels there with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
DST = os.path.join(HERE, '_ref')


def build_ref(force=False):
    if not os.path.isdir(os.path.join(REF, 'mpi4py_fft')):
        return os.path.isdir(os.path.join(DST, 'mpi4py_fft'))
    pkg = os.path.join(DST, 'mpi4py_fft')
    marker = os.path.join(DST, '.built')
    if os.path.exists(marker) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, 'mpi4py_fft'), pkg)
    # compile the one numpy-only Cython module in the staged copy
    pyx = os.path.join(pkg, 'fftw', 'utilities.pyx')
    setup = os.path.join(DST, '_setup_utilities.py')
    with open(setup, 'w') as f:
        f.write("from setuptools import setup, Extension\n"
                "from Cython.Build import cythonize\n"
                "import numpy\n"
                "setup(ext_modules=cythonize([Extension('mpi4py_fft.fftw.utilities', [%r],\n"
                "      include_dirs=[numpy.get_include()])], language_level=3))\n" % pyx)
    r = subprocess.run([sys.executable, setup, 'build_ext', '--inplace'], cwd=DST,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building reference utilities.pyx failed:\n" + r.stdout + r.stderr)
    with open(marker, 'w') as f:
        f.write('ok\n')
    return True


if __name__ == '__main__':
    print('staged' if build_ref(force='--force' in sys.argv) else 'reference not available')
