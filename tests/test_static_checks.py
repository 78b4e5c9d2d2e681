"""Static guards for the Python files that only run on the GPU box (bench.py's GPU arm, the NCCL
workers, smoke()): they compile, and every global name they reference is defined — a typo there
would otherwise first show at round end, where it costs the round's measurement."""
import builtins
import os
import symtable

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bench.py", "__graft_entry__.py", "dna-kmeres-parallel_b200/kmerb200/__init__.py",
         "dna-kmeres-parallel_b200/kmerb200/distributed.py", "tests/_nccl_worker.py", "tests/_nccl_radix_worker.py", "tests/_nccl_perseq_worker.py",
         "tests/_gloo_worker.py", "tests/_gloo_radix_worker.py", "tests/test_gpu_parity.py", "tests/test_zzz_first_gpu_run.py", "tests/_first_gpu_run_cases.py", "tests/test_gpu_parity_variants.py",
         "tests/test_multi_gpu.py", "tests/test_zz_driver_cli.py", "tools/nccl_reduce_bench.py", "tools/sanitize_smoke.py", "tests/emu/bench_dryrun.py", "tests/emu/run_under_shim.py"]


@pytest.mark.parametrize("rel", FILES)
def test_no_undefined_globals(rel):
    path = os.path.join(ROOT, rel)
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")  # also a syntax check
    module_names = set(top.get_identifiers())
    bad = []

    def walk(t):
        for s in t.get_symbols():
            if s.is_global() and s.is_referenced() and not s.is_assigned():
                n = s.get_name()
                if n not in module_names and not hasattr(builtins, n):
                    bad.append((t.get_name(), n))
        for c in t.get_children():
            walk(c)

    walk(top)
    assert not bad, bad


def test_first_run_case_list_is_complete():
    """tests/test_zzz_first_gpu_run.py runs the cases of tests/_first_gpu_run_cases.py by name, one process each"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("zzz_first", os.path.join(ROOT, "tests", "test_zzz_first_gpu_run.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert sorted(m.CASES) == sorted(m.case_names())
    assert len(set(m.CASES)) == len(m.CASES)
    # -k matches substrings: no case name may be contained in another
    assert not [(a, b) for a in m.CASES for b in m.CASES if a != b and a in b]
