"""ctypes wrapper of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
It reuses the POD struct mirrors of the product's C-ABI (the oracle shares include/glabc.h).
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
abi = importlib.import_module("gl-abc-mcmc_b200._abi")


def build(force=False):
    src = os.path.join(HERE, "glabc_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "glabc.h")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        for name in ("oracle_run_global", "oracle_run_isir", "oracle_run_mala"):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(abi.ModelPOD), C.POINTER(abi.DistPOD), C.POINTER(abi.DistPOD), C.POINTER(abi.RunPOD)]
        _lib.oracle_run_aglmcmc.restype = C.c_int
        _lib.oracle_run_aglmcmc.argtypes = [C.POINTER(abi.ModelPOD), C.POINTER(abi.DistPOD), C.POINTER(abi.DistPOD),
                                            C.POINTER(abi.RunPOD), C.POINTER(abi.AglmcmcPOD)]
        _lib.oracle_kde_fit.restype = C.c_int
        _lib.oracle_kde_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.oracle_kde_log_prob.restype = C.c_int
        _lib.oracle_kde_log_prob.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.oracle_esjd.restype = C.c_int
        _lib.oracle_esjd.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
        _lib.oracle_philox4x32_10.restype = None
        _lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32 * 4), C.POINTER(C.c_uint32 * 2), C.POINTER(C.c_uint32 * 4)]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) for x in key])
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(C.byref(c), C.byref(k), C.byref(o))
    return [int(x) for x in o]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(sampler, model, d1, d2, *, theta, y, n_steps, gf, step_base=0, chain_id_base=0, seed=0,
        rng_mode=abi.RNG_NATIVE, trace_layout=abi.TRACE_TIME_MAJOR, trace=None, trace_rows=None,
        trace_chains=None, trace_chain_off=0, write_row0=True, stats=None, aux=None, tape32=None,
        tape64=None, debug=None, K=0, threads=0, num_grad=0, tau=0.0, state64=None, tape_grad0=None, debug64=None,
        ag=None):
    """Run `sampler` ('global' | 'isir' | 'mala') on numpy buffers; theta/y/aux/stats are updated in place.

    Returns the trace ([rows, C, d] time-major or [C, rows, d] chain-major) or None."""
    L = lib()
    L.oracle_set_num_threads(int(threads))
    Cn, d = theta.shape
    for a in (theta, y, aux, stats, tape32, tape64, debug, trace, state64, tape_grad0, debug64):
        assert a is None or a.flags["C_CONTIGUOUS"]
    rows = trace_rows if trace_rows is not None else step_base + n_steps + 1
    tchains = trace_chains if trace_chains is not None else Cn
    if trace is None and trace_layout != abi.TRACE_NONE:
        shape = (rows, tchains, d) if trace_layout == abi.TRACE_TIME_MAJOR else (tchains, rows, d)
        trace = np.zeros(shape, np.float32)
    r = abi.RunPOD(n_chains=Cn, n_steps=n_steps, step_base=step_base, chain_id_base=chain_id_base, seed=seed,
                   global_frequency=float(gf), rng_mode=rng_mode, arith_mode=abi.ARITH_STRICT,
                   trace_layout=trace_layout, write_row0=int(write_row0), n_candidates=K,
                   num_grad=int(num_grad), tau=float(tau), tau64=float(tau), state64=_ptr(state64), tape_grad0=_ptr(tape_grad0),
                   debug64=_ptr(debug64),
                   trace_rows=rows, trace_chains=tchains, trace_chain_off=trace_chain_off,
                   theta=_ptr(theta), y=_ptr(y), aux=_ptr(aux), trace=_ptr(trace), stats=_ptr(stats),
                   tape32=_ptr(tape32), tape64=_ptr(tape64), debug=_ptr(debug))
    if sampler == "aglmcmc":
        st = L.oracle_run_aglmcmc(C.byref(model), C.byref(d1), C.byref(d2), C.byref(r), C.byref(ag))
        if st != 0:
            raise RuntimeError(f"oracle_aglmcmc failed with status {st}")
        return trace
    fn = {"global": L.oracle_run_global, "isir": L.oracle_run_isir, "mala": L.oracle_run_mala}[sampler]
    if sampler == "mala" and d1 is None:
        d1 = d2
    st = fn(C.byref(model), C.byref(d1), C.byref(d2), C.byref(r))
    if st != 0:
        raise RuntimeError(f"oracle_{sampler} failed with status {st}")
    return trace


def esjd(trace, layout):
    rows, chains = (trace.shape[0], trace.shape[1]) if layout == abi.TRACE_TIME_MAJOR else (trace.shape[1], trace.shape[0])
    out = np.zeros(chains, np.float32)
    st = lib().oracle_esjd(_ptr(trace), layout, rows, chains, trace.shape[2], _ptr(out))
    if st != 0:
        raise RuntimeError(f"oracle_esjd failed with status {st}")
    return out


def aglmcmc_params(*, S, alpha, hat_eps_T, rule=abi.BW_SILVERMAN, init=1, init_p=None, init_s=None, ad_idx=None, ad_noise=None,
                   ad_sim=None, ad_rec=None, ad_blk=None, init_w=None, ptr=_ptr):
    """glabc_aglmcmc_t from numpy arrays (ptr=_ptr) — tests pass their own `ptr` for device tensors"""
    rounds = 0 if ad_idx is None else ad_idx.shape[0]
    dump = 0 if ad_rec is None else ad_rec.shape[0]
    if ad_blk is not None:
        dump = ad_blk.shape[0] if dump == 0 else min(dump, ad_blk.shape[0])
    return abi.AglmcmcPOD(step_size=S, init=init, alpha=alpha, hat_eps_T=hat_eps_T, kde_rule=rule, tape_rounds=rounds,
                          init_p=ptr(init_p), init_s=ptr(init_s), ad_idx=ptr(ad_idx), ad_noise=ptr(ad_noise), ad_sim=ptr(ad_sim),
                          ad_rec=ptr(ad_rec), ad_blk=ptr(ad_blk), init_w=ptr(init_w), dump_rounds=dump)


def kde_fit(X, w, rule=abi.BW_SILVERMAN):
    n, d = X.shape
    weights, bw = np.zeros(n, np.float32), np.zeros(d, np.float32)
    st = lib().oracle_kde_fit(_ptr(X), None if w is None else _ptr(w), n, d, rule, _ptr(weights), _ptr(bw))
    assert st == 0
    return weights, bw


def kde_log_prob(X, weights, bw, x):
    n, d = X.shape
    out = np.zeros(x.shape[0], np.float32)
    st = lib().oracle_kde_log_prob(_ptr(X), _ptr(weights), _ptr(bw), n, d, _ptr(x), x.shape[0], _ptr(out))
    assert st == 0
    return out
