"""CPU-side checks of the C-ABI boundary: libglabc.so loads without a GPU, exports every symbol
include/glabc.h declares, the ctypes struct mirrors have the C layout, and the product refuses to
run (loudly) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from helpers import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "glabc.h")


def declared_symbols():
    src = open(HEADER).read()
    return re.findall(r"GLABC_API\s+[\w\s\*]+?\b(glabc_\w+)\s*\(", src)


def test_library_loads_and_exports_every_declared_symbol():
    lib = abi.load()
    names = declared_symbols()
    assert len(names) >= 12 and "glabc_run_global" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in glabc.h but not exported"
    assert set(names) == set(abi._SIGNATURES), "ctypes signature table out of sync with glabc.h"
    assert lib.glabc_version() == 1
    assert lib.glabc_status_string(abi.ERR_NO_DEVICE) == b"no usable CUDA device"


def test_struct_layouts_match_the_header(tmp_path):
    """compile a tiny C program that prints sizeof/offsetof and compare with ctypes"""
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "glabc.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(glabc_model_t), sizeof(glabc_dist_t), sizeof(glabc_run_t));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(glabc_run_t, tau), offsetof(glabc_run_t, trace_rows),
         offsetof(glabc_run_t, trace_row_base), offsetof(glabc_run_t, theta), offsetof(glabc_run_t, aux),
         offsetof(glabc_run_t, tape_dump), offsetof(glabc_run_t, stream));
  printf("%zu %zu\n", offsetof(glabc_model_t, eps_log_scale), offsetof(glabc_dist_t, mix_log_w));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = [int(x) for x in out]
    R = abi.RunPOD
    want = [C.sizeof(abi.ModelPOD), C.sizeof(abi.DistPOD), C.sizeof(R), R.tau.offset, R.trace_rows.offset,
            R.trace_row_base.offset, R.theta.offset, R.aux.offset, R.tape_dump.offset, R.stream.offset,
            abi.ModelPOD.eps_log_scale.offset, abi.DistPOD.mix_log_w.offset]
    assert got == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import glabc_b200 as g
    with pytest.raises(abi.GlabcError) as ei:
        abi.Context()
    assert ei.value.status == abi.ERR_NO_DEVICE
    model = g.Mixture_set(0.05)
    lp = g.DiagGaussian(2, torch.zeros(1, 2), torch.log(torch.tensor([0.35, 0.35])))
    gp = g.DiagGaussian(2, torch.zeros(2), torch.zeros(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.GlobalMCMC(model, 10, torch.zeros(2), torch.zeros(1, 2), gp, None, 0.5, lp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.esjd(torch.zeros(10, 2))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gl-abc-mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|liboracle|oracle[/.]\w", text), f"{f} uses the oracle"


def test_expand_events_on_the_host():
    """glabc_expand_events (no GPU): a GLABC_TRACE_EVENTS buffer expanded into dense chain-major rows equals the numpy expansion,
    for every row width, ragged run lengths (runs shorter than a cache line included), and bad buffers are refused untouched"""
    import numpy as np
    lib = abi.load()
    rng = np.random.default_rng(5)
    for d in (1, 2, 3, 4):
        chains, cap, T = 37, 64, 777
        ev = np.zeros((chains, cap, 1 + d), np.float32)
        want = np.zeros((chains, T, d), np.float32)
        for c in range(chains):
            m = int(rng.integers(1, cap - 1))
            rows = np.sort(rng.choice(np.arange(1, T), size=m - 1, replace=False)) if m > 1 else np.zeros(0, np.int64)
            rows = np.concatenate([[0], rows]).astype(np.uint32)
            vals = rng.standard_normal((m, d)).astype(np.float32)
            ev[c, 0, 0] = np.array([m], np.uint32).view(np.float32)[0]
            ev[c, 1:m + 1, 0] = rows.view(np.float32)
            ev[c, 1:m + 1, 1:] = vals
            ends = np.concatenate([rows[1:], [T]])
            for k in range(m):
                want[c, rows[k]:ends[k]] = vals[k]
        for threads in (1, 5):
            got = np.full((chains, T, d), np.nan, np.float32)
            st = lib.glabc_expand_events(ev.ctypes.data, chains, cap, d, 0, T - 1, got.ctypes.data, T, threads)
            assert st == abi.OK and np.array_equal(got, want), (d, threads)
        bad = ev.copy()
        bad[3, 0, 0] = np.array([cap], np.uint32).view(np.float32)[0]          # more moves than the capacity
        got = np.full((chains, T, d), np.nan, np.float32)
        assert lib.glabc_expand_events(bad.ctypes.data, chains, cap, d, 0, T - 1, got.ctypes.data, T, 2) == abi.ERR_INVALID
        assert np.isnan(got).all()
