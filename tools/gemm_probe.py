"""GPU probe for polus_gemm_tc: every major combination / tile width / epilogue against the CUDA-core
GEMM on device and numpy on host.  Each case runs in a fresh subprocess so a trap in one case does
not poison the others.  Usage: python tools/gemm_probe.py [case_index]"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, M, N, K, a_mn, b_mn, batch0, batch1, c_f32, bias, act, split_k
    ("kk_128x128x64", 128, 128, 64, 0, 0, 1, 1, 0, 0, 0, 1),
    ("kk_256x256x256", 256, 256, 256, 0, 0, 1, 1, 1, 0, 0, 1),
    ("kmn_128x128x64", 128, 128, 64, 0, 1, 1, 1, 1, 0, 0, 1),
    ("mnmn_128x128x64", 128, 128, 64, 1, 1, 1, 1, 1, 0, 0, 1),
    ("mnk_128x128x64", 128, 128, 64, 1, 0, 1, 1, 1, 0, 0, 1),
    ("kk_n64", 256, 64, 128, 0, 0, 1, 1, 1, 0, 0, 1),
    ("kmn_fwd_bias_gelu", 1024, 3072, 768, 0, 1, 1, 1, 0, 1, 1, 1),
    ("kk_dgrad", 1024, 768, 3072, 0, 0, 1, 1, 0, 0, 0, 1),
    ("mnmn_wgrad_split", 768, 3072, 4096, 1, 1, 1, 1, 1, 0, 0, 8),
    ("kk_batched_scores", 256, 256, 64, 0, 0, 12, 4, 0, 0, 0, 1),
    ("kmn_batched_pv", 256, 64, 256, 0, 1, 12, 4, 0, 0, 0, 1),
    ("mnmn_batched_dv", 256, 64, 256, 1, 1, 12, 4, 0, 0, 0, 1),
    ("ragged_kmn", 200, 136, 200, 0, 1, 1, 1, 1, 1, 3, 1),
    ("ragged_mnmn", 328, 72, 520, 1, 1, 3, 1, 1, 0, 0, 1),
    ("big_fwd", 8192, 2304, 768, 0, 1, 1, 1, 0, 1, 0, 1),
]


def run_case(idx):
    from polus_b200 import _lib, device
    name, M, N, K, a_mn, b_mn, b0, b1, c_f32, use_bias, act, split = CASES[idx]
    device.init(0)
    rng = np.random.default_rng(1234 + idx)
    nb = b0 * b1
    a_shape = (nb, K, M) if a_mn else (nb, M, K)
    b_shape = (nb, K, N) if b_mn else (nb, N, K)
    A = device.bf16_round(rng.standard_normal(a_shape, dtype=np.float32))
    B = device.bf16_round(rng.standard_normal(b_shape, dtype=np.float32) * 0.1)
    bias = rng.standard_normal(N).astype(np.float32) if use_bias else None
    dA = device.Buffer(A.size * 2); device.upload(dA.ptr, device.f32_to_bf16_bits(A))
    dB = device.Buffer(B.size * 2); device.upload(dB.ptr, device.f32_to_bf16_bits(B))
    dbias = None
    if use_bias:
        dbias = device.Buffer(N * 4); device.upload(dbias.ptr, bias)
    csz = 4 if c_f32 else 2
    dC = device.Buffer(nb * M * N * csz, zero=True)
    dR = device.Buffer(nb * M * N * 4, zero=True)
    device.synchronize()

    def mk(Cptr, c_dtype, split_k, accumulate):
        g = _lib.Gemm()
        g.M, g.N, g.K, g.batch0, g.batch1 = M, N, K, b0, b1
        g.A = _lib.Operand(dA.ptr, M if a_mn else K, M * K, M * K * b0, a_mn, _lib.BF16)
        g.B = _lib.Operand(dB.ptr, N if b_mn else K, N * K, N * K * b0, b_mn, _lib.BF16)
        g.C, g.ldc, g.cbs0, g.cbs1, g.c_dtype = Cptr, N, M * N, M * N * b0, c_dtype
        g.C2, g.bias = None, (dbias.ptr if dbias else None)
        g.alpha, g.act, g.accumulate, g.split_k = 0.5, act, accumulate, split_k
        return g

    g_tc = mk(dC.ptr, _lib.F32 if c_f32 else _lib.BF16, split, 1 if split > 1 else 0)
    g_ref = mk(dR.ptr, _lib.F32, 1, 0)
    assert _lib.call("polus_gemm_tc_supported", C.byref(g_tc)) == 1, _lib.last_error()
    _lib.call("polus_gemm_small", C.byref(g_ref), device.stream())
    _lib.call("polus_gemm_tc", C.byref(g_tc), device.stream())
    device.synchronize()
    ref = device.download(dR.ptr, (nb, M, N), np.float32)
    if c_f32:
        out = device.download(dC.ptr, (nb, M, N), np.float32)
    else:
        out = device.bf16_bits_to_f32(device.download(dC.ptr, (nb, M, N), np.uint16))
    err = float(np.abs(out - ref).max())
    scale = float(np.abs(ref).max()) + 1e-6
    res = {"case": name, "max_abs_err": err, "ref_max": scale, "rel": err / scale}
    # host check on small problems
    if M * N * K * nb <= 64 * 1024 * 1024:
        Am = np.swapaxes(A, 1, 2) if a_mn else A
        Bm = np.swapaxes(B, 1, 2) if b_mn else B
        host = 0.5 * np.einsum("bmk,bnk->bmn", Am.astype(np.float64), Bm.astype(np.float64))
        if use_bias:
            host = host + bias
        if act == 1:
            from math import erf
            host = 0.5 * host * (1 + np.vectorize(erf)(host / np.sqrt(2)))
        elif act == 3:
            host = host / (1 + np.exp(-host))
        res["host_rel"] = float(np.abs(out - host).max() / (np.abs(host).max() + 1e-6))
    # timing (10 reps)
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    _lib.call("polus_event_create", C.byref(ev0)); _lib.call("polus_event_create", C.byref(ev1))
    for _ in range(3):
        _lib.call("polus_gemm_tc", C.byref(g_tc), device.stream())
    _lib.call("polus_event_record", ev0, device.stream())
    reps = 20
    for _ in range(reps):
        _lib.call("polus_gemm_tc", C.byref(g_tc), device.stream())
    _lib.call("polus_event_record", ev1, device.stream())
    device.synchronize()
    ms = C.c_float()
    _lib.call("polus_event_elapsed_ms", ev0, ev1, C.byref(ms))
    res["us"] = ms.value * 1000 / reps
    res["tflops"] = 2.0 * M * N * K * nb / (ms.value / reps * 1e-3) / 1e12
    tol = 2e-2 if not c_f32 else 2e-3
    res["ok"] = bool(res["rel"] < tol)
    print(json.dumps(res), flush=True)
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(run_case(int(sys.argv[1])))
    bad = 0
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
            if r.returncode != 0:
                bad += 1
                print(f"FAIL[{c[0]}] rc={r.returncode} {out} :: {r.stderr.strip()[-400:]}", flush=True)
            else:
                print(out, flush=True)
        except subprocess.TimeoutExpired:
            bad += 1
            print(f"TIMEOUT[{c[0]}]", flush=True)
    print(f"gemm_probe: {len(CASES) - bad}/{len(CASES)} ok")
    sys.exit(1 if bad else 0)
