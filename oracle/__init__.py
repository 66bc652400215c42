"""CPU oracle for the energy+gradient hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``membrane_solver_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU baseline, never as the shipped path.

Contents
--------
``ref_modules``  NumPy restatement of the reference's module-level algorithm
                 (surface, volume, bending, tilt, bending_tilt) on plain arrays.
``ref_kernels.c`` Plain-C restatement of the five ``fortran_kernels/*.f90``
                 subroutines (gfortran is not in this image, so the Fortran
                 itself cannot be compiled).
``ckernels``     ctypes binding of the C restatement, exposing the same
                 ``KernelSpec(func, expects_transpose)`` seam as
                 ``fortran_kernels/loader.py:15-20``.

Parity pinning: ``tests/golden/*.npz`` are produced by importing the real
reference (``tests/golden/generate_golden.py``, run in the build container
where ``/root/reference`` exists) and ``tests/test_oracle_golden.py`` checks
this oracle against every one of them.
"""
