"""Run a GPU-only Python script of this repo on a box without a GPU — TEST INFRASTRUCTURE ONLY:
the stand-ins of bench_dryrun.py (emulator library in the place of libkmerb200.so, torch's CUDA
surface on the CPU, NCCL -> gloo) are installed, then the script runs as __main__.

    python tests/emu/run_under_shim.py tests/_nccl_radix_worker.py
    python tests/emu/run_under_shim.py -m pytest tests/_first_gpu_run_cases.py -m gpu -q      (with KC_FIRST_RUN_SCALE=0.002)
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_dryrun  # noqa: E402

if __name__ == "__main__":
    bench_dryrun.install_shim()
    if sys.argv[1] == "-m":   # python tests/emu/run_under_shim.py -m pytest tests/_first_gpu_run_cases.py -m gpu ...
        sys.argv = sys.argv[2:]
        runpy.run_module(sys.argv[0], run_name="__main__", alter_sys=True)
    else:
        script = sys.argv[1]
        sys.argv = sys.argv[1:]
        runpy.run_path(script, run_name="__main__")
