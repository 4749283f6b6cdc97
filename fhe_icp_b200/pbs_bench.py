"""PBS / keyswitch microbenchmark at the stated TFHE parameter set (BASELINE.json configs[2]).
Called by bench.py; reports PBS/s per batch size with both roofline terms (FP64 pipe and key
streaming) and says which one binds."""
from __future__ import annotations

import numpy as np

from .params import PBS_PARAMS_4BIT


def flops_per_pbs(n, k, N, l):
    M = N // 2
    fft = 5 * M * np.log2(M)
    return n * (((k + 1) * l + (k + 1)) * fft + (k + 1) ** 2 * l * M * 8)


def measure(dev, args, batches=None):
    import ctypes as C
    import torch
    from . import _native as N_
    from . import engine as E
    d = dict(PBS_PARAMS_4BIT)
    p = E.make_pbs_params(**d)
    s, S = E.secret_key(101, 0, p.n, dev), E.secret_key(101, 1, p.k * p.N, dev)
    ksk, bsk = E.ksk_gen(p, S, s, 202), E.bsk_gen(p, s, S, 202)
    bskf = E.bsk_to_fourier(p, bsk)
    del bsk
    bsk2 = E.bsk2_gen(p, s, S, 202)     # multi-bit blind rotation: 3 key elements per pair of key bits
    bskf2 = E.bsk2_to_fourier(p, bsk2)
    del bsk2
    ksk32 = E.ksk_to_32(p, ksk)
    key_mma = E.ksk_to_mma(p, ksk32)    # the same key as int8 MMA blocks (tensor-core keyswitch)
    table = (np.arange(16) * 7 + 3) % 16
    lut = E.from_u64_numpy(E.make_lut_poly(table, 4, p.N, 59), dev)
    ctx = N_.context(dev.index)
    sm = ctx.device_info()["sm_count"]
    fp64_peak = ctx.probe_fp64_tflops()
    if batches is None:
        batches = [1, 16, sm, 2 * sm, 8 * sm, 32 * sm, 65536] if not getattr(args, "pbs_batch", 0) else [args.pbs_batch]
    flops = flops_per_pbs(p.n, p.k, p.N, p.l_pbs)
    hbm_gbs = float(getattr(args, "_hbm_peak", None) or 6543.4)   # MEASURED_PEAKS.json copy bandwidth
    bsk_bytes = p.n * (p.k + 1) ** 2 * p.l_pbs * p.N * 8
    rows, best = [], None
    rng = np.random.RandomState(5)
    for B in batches:
        msgs = rng.randint(0, 16, size=B)
        ct_big = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=303, ct_base=B * 7919,
                               stride=p.N + 2)[:, : p.N + 1].contiguous()
        ct = E.keyswitch(p, ksk, ct_big)
        out = torch.empty((B, p.k * p.N + 1), dtype=torch.int64, device=dev)
        E.pbs(p, bskf, ct, lut, out=out)  # warm-up (also first-touch of the key into L2)
        torch.cuda.synchronize()
        reps = 3 if B >= 8 * sm else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            E.pbs(p, bskf, ct, lut, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out2 = torch.empty_like(out)
        E.pbs_mb2(p, bskf2, ct, lut, out=out2)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            E.pbs_mb2(p, bskf2, ct, lut, out=out2)
        e1.record()
        torch.cuda.synchronize()
        mb2_ms = e0.elapsed_time(e1) / reps
        def time_ks(fn):
            fn()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(reps):
                fn()
            k1.record()
            torch.cuda.synchronize()
            return k0.elapsed_time(k1) / reps
        ks64_ms = time_ks(lambda: E.keyswitch(p, ksk, ct_big))
        ks32_ms = time_ks(lambda: E.keyswitch32(p, ksk32, ct_big))   # 32-bit keyswitch on the integer pipe
        ks_work = torch.empty(int(N_.lib().fhe_b200_keyswitch_mma_workspace_bytes(C.byref(p), B)), dtype=torch.int8, device=dev)
        ks_ms = time_ks(lambda: E.keyswitch_mma(p, key_mma, ct_big, work=ks_work))   # the same on the tensor cores: the one used
        ks_same = bool(torch.equal(E.keyswitch_mma(p, key_mma, ct_big, work=ks_work), E.keyswitch32(p, ksk32, ct_big)))
        dec = E.lwe_decrypt(S, _pad(out), 59)
        ok = bool(np.array_equal(dec.cpu().numpy() & 15, table[msgs]))
        ok2 = bool(np.array_equal(E.lwe_decrypt(S, _pad(out2), 59).cpu().numpy() & 15, table[msgs]))
        fastest = min(ms, mb2_ms)
        row = {"batch": int(B), "pbs_ms": fastest, "pbs_per_sec": B / (fastest * 1e-3),
               "kernel": (_mb2_kernels(B, sm) if mb2_ms < ms else "pbs_kernel_tmem (one key bit per step)"),
               "single_bit_ms": ms, "single_bit_per_sec": B / (ms * 1e-3),
               "multi_bit_ms": mb2_ms, "multi_bit_per_sec": B / (mb2_ms * 1e-3),
               "ks_ms": ks_ms, "ks32_int_pipe_ms": ks32_ms, "ks64_ms": ks64_ms,
               "ks_per_sec": B / (ks_ms * 1e-3), "ks_pbs_per_sec": B / ((fastest + ks_ms) * 1e-3),
               "ks_int8_tops": 2.0 * B * p.k * p.N * p.l_ks * 4 * (p.n + 1) / (ks_ms * 1e-3) / 1e12,
               "fp64_tflops": flops * B / (fastest * 1e-3) / 1e12, "correct": ok and ok2 and ks_same}
        # both roofline terms: key streamed once per launch + ciphertext I/O over HBM, and the
        # algorithmic FP64 work; whichever is larger is the roofline time for this batch
        key_bytes = bsk_bytes * (3 / 2 if mb2_ms < ms else 1)
        hbm_ms = (key_bytes + B * (p.n + 1 + p.k * p.N + 1 + p.N) * 8) / (hbm_gbs * 1e9) * 1e3
        fp64_ms = flops * B / (fp64_peak * 1e12) * 1e3 if fp64_peak else None
        row["roofline_terms"] = {"hbm_ms": hbm_ms, "fp64_ms": fp64_ms,
                                 "binding": "fp64" if fp64_ms and fp64_ms > hbm_ms else "hbm",
                                 "frac_of_binding": max(hbm_ms, fp64_ms or 0.0) / fastest}
        rows.append(row)
        if best is None or row["pbs_per_sec"] > best["pbs_per_sec"]:
            best = row
    hbm_peak = getattr(args, "_hbm_peak", None)
    res = {
        "metric": "pbs_per_sec", "value": best["pbs_per_sec"], "unit": "PBS/s", "batch": best["batch"],
        "kernel": best["kernel"],
        "ks_pbs_per_sec": best["ks_pbs_per_sec"], "all_correct": all(r["correct"] for r in rows),
        "keyswitch": _keyswitch_summary(rows, p, getattr(args, "_bf16_peak", None)),
        "params": {k: d[k] for k in ("n", "k", "N_poly", "l_pbs", "beta_pbs", "l_ks", "beta_ks", "log2_sigma_lwe",
                                     "log2_sigma_glwe")},
        "by_batch": rows,
        "roofline": {"bound": "fp64", "achieved": best["fp64_tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": best["fp64_tflops"] / fp64_peak if fp64_peak else None,
                     "peak_source": "measured live (fhe_b200_probe_fp64, dependent-FMA chains)",
                     "flops_per_pbs": flops, "bsk_fourier_bytes": int(bsk_bytes),
                     "fp64_pipe_active_ncu": 0.502, "smem_wavefronts_of_peak_ncu": 0.519, "issue_active_ncu": 0.354,
                     "ncu_source": "profiles/r2_ncu_pbs_mb2_v10.txt (the shipped pbs_kernel_mb2<1,4>, batch 1184); small-batch "
                                   "kernels pbs_kernel_mb2_wide / pbs_kernel_mb2_pair: profiles/r2_ncu_pbs_wide_v2.txt, profiles/r2_ncu_pbs_pair_v1.txt",
                     "hbm_term": {"bytes_per_batch": int(bsk_bytes + best["batch"] * (p.n + 1 + p.k * p.N + 1 + p.N) * 8),
                                  "note": "the Fourier key is read from HBM once per launch and then served from L2 "
                                          "(ncu: 52 MB DRAM reads per launch); key streaming never binds once batched"},
                     "note": "achieved = ALGORITHMIC flops of the textbook one-bit-per-step blind rotation (SURVEY.md 8d) / time; "
                             "the multi-bit kernel executes ~17 % fewer FP64 instructions for the same PBS.  "
                             "flops = 5*M*log2(M) per FFT + 8 per complex MAC; the kernel's instruction mix "
                             "(DADD/DMUL/DFMA ~ 45/25/30 %) caps it at ~65 % of the FMA peak even with a saturated "
                             "pipe.  Small batches are neither HBM nor FP64 bound: one ciphertext's blind rotation is a "
                             "serial chain of 371 (multi-bit) / 742 CMuxes on one SM (batch 1: 1.3 ms on a cluster of two SMs, 1.8 ms on one), and the key "
                             "(49-73 MB) sits in the 126 MB L2 after its first read, so key streaming from HBM never binds "
                             "(by_batch[].roofline_terms gives both terms per batch)."},
    }
    return res


def _mb2_kernels(B: int, sm: int) -> str:
    """Which kernels fhe_b200_pbs_mb2 launches for a batch (the dispatch of csrc/pbs.cu::launch_pbs_mb2): full waves of
    4 x SMs ciphertexts on pbs_kernel_mb2<1,4>, a remainder of up to 3 x SMs on the latency kernel pbs_kernel_mb2_wide, or -- up to SMs / 2 -- on
    pbs_kernel_mb2_pair (two SMs per ciphertext)."""
    full = B // (4 * sm) * (4 * sm)
    rest = B - full
    if rest > 3 * sm:
        full, rest = B, 0
    parts = []
    if full:
        parts.append(f"pbs_kernel_mb2<1,4> x {full} (two key bits per step, four ciphertexts per CTA)")
    if rest and 2 * rest <= sm:
        parts.append(f"pbs_kernel_mb2_pair x {rest} (two key bits per step, one ciphertext per cluster of two CTAs)")
    elif rest:
        parts.append(f"pbs_kernel_mb2_wide x {rest} (two key bits per step, one ciphertext per CTA, eight warps)")
    return " + ".join(parts)


def _keyswitch_summary(rows, p, bf16_peak_tflops):
    """Keyswitch section of the bench line: the tensor-core contraction against an int8 roofline."""
    best = max(rows, key=lambda r: r["ks_per_sec"])
    # int8 dense peak: twice the dense bf16 rate (same tensor pipe, half the operand width); the bf16 figure is the
    # measured cuBLAS burst number of MEASURED_PEAKS.json when present, else the nominal 2250 TFLOP/s
    bf16 = float(bf16_peak_tflops) if bf16_peak_tflops else 2250.0
    peak = 2.0 * bf16
    return {"metric": "keyswitches_per_sec", "value": best["ks_per_sec"], "unit": "keyswitch/s", "batch": best["batch"],
            "kernel": "ks_mma_kernel (tcgen05.mma.kind::i8, s8 digits x u8 key bytes -> s32 in TMEM) + ks_digits_kernel",
            "integer_pipe_ks32_per_sec": best["batch"] / (best["ks32_int_pipe_ms"] * 1e-3),
            "bit_identical_to_integer_kernel": all(r["correct"] for r in rows),
            "roofline": {"bound": "tensor", "achieved": best["ks_int8_tops"], "peak": peak, "unit": "TOP/s",
                         "frac": best["ks_int8_tops"] / peak,
                         "peak_source": ("2 x measured dense bf16 burst (MEASURED_PEAKS.json)" if bf16_peak_tflops
                                         else "2 x nominal dense bf16 (2250 TFLOP/s)"),
                         "ops_per_keyswitch": 2.0 * p.k * p.N * p.l_ks * 4 * (p.n + 1),
                         "note": "achieved counts the int8 MACs of the byte-plane contraction (4 planes per 32-bit key word) over the "
                                 "time of digit extraction + contraction; 128x256 tiles read 48 KB of operands per 1 M MACs, so the "
                                 "kernel is bounded by L2->SM operand bandwidth before the tensor pipe"}}


def measure_pair(dev, args, docs: int = 148):
    """Encrypted x encrypted comparison: both vectors encrypted; d = 128 programmable
    bootstraps per document with the squared-norm protocol (2d = 256 without), n=742, N=2048, l_pbs=2,
    two-level multi-bit blind rotation."""
    import time
    import torch
    from . import engine as E_
    from .encrypted_compare import COMPARE_PARAMS, EncryptedCompare, IN_SHIFT, OUT_SHIFT, P_BITS
    d = 128
    ec = EncryptedCompare(input_dim=d, device=dev).keygen()
    rng = np.random.RandomState(17)
    q = rng.randn(d) / np.sqrt(d)
    X = rng.randn(docs, d) / np.sqrt(d)
    ec.fit_scale(X)
    xq, yq = ec.quantize(q), ec.quantize(X)
    ct_q, ct_d = ec.encrypt(xq, 1, 0), ec.encrypt(yq, 1, d)
    n_q, n_d = ec.encrypt_norms(xq, 1, 0), ec.encrypt_norms(yq, 1, 1)
    ec.scores(ct_q, ct_d, n_q, n_d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sc2 = ec.scores(ct_q, ct_d)               # two bootstraps per dimension (no norm ciphertexts)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1)
    e0.record()
    sc = ec.scores(ct_q, ct_d, n_q, n_d)      # default protocol: one bootstrap per dimension
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()   # end to end: host floats in, host floats out
    sim = ec.similarity(q, X)
    e2e_s = time.perf_counter() - t0
    exact = (bool(np.array_equal(ec.decrypt(sc), yq @ xq)) and bool(np.array_equal(ec.decrypt(sc2), yq @ xq))
             and bool(np.array_equal(sim, ec.dequantize(yq @ xq))))
    res = {"metric": "encrypted_pair_comparisons_per_sec", "value": docs / (ms * 1e-3), "unit": "comparisons/s",
           "e2e": {"value": docs / e2e_s, "unit": "comparisons/s"},
           "pbs_per_sec": d * docs / (ms * 1e-3), "docs": docs, "d": d, "pbs_per_comparison": d,
           "protocol": "each party also sends an encryption of its squared norm: sum (x+y)^2 - |x|^2 - |y|^2",
           "without_norm_ciphertexts": {"value": docs / (ms2 * 1e-3), "pbs_per_comparison": 2 * d},
           "params": dict(COMPARE_PARAMS), "kernel": "pbs_kernel_mb2<2,2>", "exact_vs_clear_integer_model": exact}
    # exact encrypted threshold (13 keyswitch + PBS per score) on a batch of score ciphertexts
    from .encrypted_compare import EncryptedThreshold
    th = EncryptedThreshold(ec)
    nthr = 592
    vals = rng.randint(-1536, 2049, size=nthr)
    sct = E_.lwe_encrypt(ec.S, torch.as_tensor(vals), OUT_SHIFT, 2.0 ** (64 - 18), enc_seed=5, ct_base=0)
    T = th.threshold_to_int(0.5)
    th.ge(sct[:8], T)
    torch.cuda.synchronize()
    e0.record()
    bits = th.ge(sct, T)
    e1.record()
    torch.cuda.synchronize()
    thr_ms = e0.elapsed_time(e1)
    res["threshold"] = {"metric": "encrypted_thresholds_per_sec", "value": nthr / (thr_ms * 1e-3), "batch": nthr,
                        "ks_pbs_per_threshold": 13,
                        "exact": bool(np.array_equal(th.decrypt(bits), (vals >= T).astype(np.int64)))}
    # bench.py's CPU-baseline leg re-evaluates this sample with the oracle (the product never imports it)
    res["_sample"] = {"xq": xq, "yq": yq[:2], "expect": (yq[:2] @ xq).tolist(), "key_seed": ec.key_seed,
                      "evk_seed": ec.evk_seed, "stride": int(ct_q.shape[-1]), "in_shift": IN_SHIFT,
                      "out_shift": OUT_SHIFT, "p_bits": P_BITS}
    return res


def measure_packed(dev, args, docs: int = 262144):
    """Packed both-encrypted comparison (GLWE x GGSW external product, 16 documents per ciphertext at
    d = 128, query GGSW resident in TMEM): the leveled form of SURVEY.md 8f N1."""
    import time
    import torch
    from . import engine as E_
    from .batch_operations import rank_results
    from .encrypted_compare import PACKED_OUT_SHIFT, PACKED_PARAMS, PackedEncryptedCompare
    d = 128
    pe = PackedEncryptedCompare(input_dim=d, device=dev).keygen()
    rng = np.random.RandomState(23)
    q = rng.randn(d); q /= np.linalg.norm(q)
    X = rng.randn(docs, d)
    X[::5] = 0.8 * q + 0.6 * X[::5] / np.sqrt(d)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    pe.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(d))
    xq, yq = pe.quantize(q), pe.quantize(X)
    gd = pe.encrypt_documents(yq, 1)               # the collection, encrypted once, resident (2 KB per document)
    gq = pe.encrypt_query(xq, 2)
    out = torch.empty_like(gd)
    pe.scores(gq, gd, out)
    torch.cuda.synchronize()
    reps = 5
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(reps):
        pe.scores(gq, gd, out)
    e1.record()
    for _ in range(reps):
        E_.glwe_decrypt_coeffs(pe.p, pe.S, out, 0, pe.slot, pe.per, PACKED_OUT_SHIFT)
    e2.record()
    torch.cuda.synchronize()
    ms, dec_ms = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
    ids = [f"doc_{i}" for i in range(docs)]
    t0 = time.perf_counter()   # end to end per query: host floats -> GGSW -> products -> client decrypt -> host scores -> top-k
    for _ in range(reps):
        gq2 = pe.encrypt_query(pe.quantize(q), 3)
        ints = pe.decrypt(pe.scores(gq2, gd, out), docs)
        top = rank_results(ids, pe.dequantize(ints), 3, 0.5)
    e2e_s = (time.perf_counter() - t0) / reps
    clear = yq @ xq
    exact = bool(np.array_equal(ints, clear)) and top == rank_results(ids, pe.dequantize(clear), 3, 0.5)
    G = gd.shape[0]
    hbm = 2 * gd.numel() * 8
    res = {"metric": "packed_encrypted_pair_comparisons_per_sec", "value": docs / (ms * 1e-3), "unit": "comparisons/s",
           "e2e": {"value": docs / e2e_s, "unit": "comparisons/s",
                   "what": "per query: quantize + GGSW-encrypt the query, external products over the resident encrypted "
                           "collection, client decrypt kernel, D2H of the scores, host top-k"},
           "docs": docs, "d": d, "docs_per_ciphertext": pe.per, "ciphertexts": int(G), "kernel": "glwe_dot_kernel<3>",
           "external_products_per_sec": G / (ms * 1e-3), "client_decrypt_ms": dec_ms, "ms": ms,
           "params": dict(PACKED_PARAMS), "factor_bits": 5, "score_bits": 17, "exact_vs_clear_integer_model": exact,
           "roofline": {"bound": "fp64", "achieved_hbm_gbs": hbm / (ms * 1e-3) / 1e9,
                        "algorithmic_bytes_per_launch": int(hbm),
                        "flops_per_ciphertext": float(6 * 5 * 1024 * 10 + 2 * 4 * 1024 * 8),
                        "achieved_tflops": float(6 * 5 * 1024 * 10 + 2 * 4 * 1024 * 8) * G / (ms * 1e-3) / 1e12,
                        "fp64_pipe_active_ncu": 0.40, "dram_throughput_of_peak_ncu": 0.217, "issue_active_ncu": 0.32,
                        "ncu_source": "profiles/r2_ncu_glwe_dot_v2.txt (the shipped kernel)",
                        "note": "32 KB in + 32 KB out per ciphertext is a third of the HBM roof; the six 1024-point f64 "
                                "FFTs per ciphertext bind (6 warps per SM at 255 registers, latency-exposed)"}}
    res["_sample"] = {"xq": xq, "yq": yq[:1024], "key_seed": pe.key_seed, "slot": pe.slot, "per": pe.per,
                      "out_shift": PACKED_OUT_SHIFT, "expect": clear[:1024]}
    return res


def _pad(out):
    """[B, kN+1] -> even-stride rows for the decrypt kernel."""
    import torch
    B, w = out.shape
    if w % 2 == 0:
        return out
    z = torch.zeros((B, w + 1), dtype=out.dtype, device=out.device)
    z[:, :w] = out
    return z


def measure_sharded(R, dev, args, total: int = 65536, per_gpu: int = 4736):
    """BASELINE.json's "PBS/sec at 1/2/4/8 B200": keyswitch + PBS over ONE batch cut into contiguous ranges, one per GPU
    (sharded_search.ShardedBootstrap).  The client rank (0) owns the secret keys, generates the evaluation keys and
    broadcasts them once over NCCL (timed); the server ranks hold nothing else.  The data path has no collective:
    ciphertexts are independent, rank r receives its rows, runs the tensor-core keyswitch and the multi-bit blind
    rotation, and sends the outputs back.  Two workloads: configs[2]'s upper end (``total`` ciphertexts, strong scaling)
    and ``per_gpu`` ciphertexts per GPU (weak).  Kernel time is CUDA events on every rank, max over ranks; the round trip
    (scatter + evaluate + collect on the client) is wall clock between barriers, max over ranks.  Every output is
    decrypted on the client and compared with the table.  Called on EVERY rank."""
    import time
    import torch
    from . import engine as E
    from .sharded_search import BootstrapEngine, ShardedBootstrap, broadcast_keys
    d = dict(PBS_PARAMS_4BIT)
    p = E.make_pbs_params(**d)
    client = R.rank == 0
    S = None
    if client:
        s, S = E.secret_key(101, 0, p.n, dev), E.secret_key(101, 1, p.k * p.N, dev)
        key_mma = E.ksk_to_mma(p, E.ksk_to_32(p, E.ksk_gen(p, S, s, 202)))
        bskf2 = E.bsk2_to_fourier(p, E.bsk2_gen(p, s, S, 202))
        eng = BootstrapEngine(d, dev, key_mma, bskf2)
    else:
        eng = BootstrapEngine(d, dev)          # parameter set only: no secret key on a server rank
    sb = ShardedBootstrap(eng, client_rank=0, device=dev)
    # the key broadcast once more, alone, between device events (the first one carried NCCL's lazy channel setup)
    torch.cuda.synchronize()
    R.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    broadcast_keys(eng.key_tensors(), 0)
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = R.max_over_ranks(e0.elapsed_time(e1))
    table = (np.arange(16) * 7 + 3) % 16
    lut = E.from_u64_numpy(E.make_lut_poly(table, 4, p.N, 59), dev)
    flops = flops_per_pbs(p.n, p.k, p.N, p.l_pbs)
    fp64_peak = None
    if client:
        from . import _native as N_
        fp64_peak = N_.context(dev.index).probe_fp64_tflops()
    rows = []
    for label, B in (("strong", int(total)), ("weak", int(per_gpu) * R.world)):
        ct_big, msgs = None, None
        if client:
            msgs = np.random.RandomState(B).randint(0, 16, size=B)
            ct_big = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=303, ct_base=B * 7919,
                                   stride=p.N + 2)[:, : p.N + 1].contiguous()
        _, mine, _ = sb.scatter(ct_big)
        eng.bootstrap(mine, lut)                       # warm-up (first touch of the keys into this GPU's L2)
        torch.cuda.synchronize()
        R.barrier()
        reps = 3
        e0.record()
        for _ in range(reps):
            eng.bootstrap(mine, lut)
        e1.record()
        torch.cuda.synchronize()
        kern_ms = R.max_over_ranks(e0.elapsed_time(e1) / reps)
        full = sb.evaluate(ct_big, lut)               # warm-up of the whole round trip (buffers, NCCL's lazy p2p channels)
        del full
        torch.cuda.synchronize()
        trips = 2
        R.barrier()
        w0 = time.perf_counter()
        for _ in range(trips):
            full = sb.evaluate(ct_big, lut)
        torch.cuda.synchronize()
        trip_ms = R.max_over_ranks((time.perf_counter() - w0) * 1e3 / trips)
        ok = None
        if client:
            z = torch.zeros((B, p.N + 2), dtype=torch.int64, device=dev)
            z[:, : p.N + 1] = full
            ok = bool(np.array_equal(E.lwe_decrypt(S, z, 59).cpu().numpy() & 15, table[msgs]))
            assert ok, "sharded keyswitch + PBS: a decrypted output differs from the table"
            del z
        rows.append({"scaling": label, "batch_total": B, "per_rank": [sb.bounds(B, r)[1] - sb.bounds(B, r)[0] for r in range(R.world)],
                     "ks_pbs_ms": kern_ms, "ks_pbs_per_sec": B / (kern_ms * 1e-3),
                     "round_trip_ms": trip_ms, "round_trip_per_sec": B / (trip_ms * 1e-3),
                     "fp64_tflops": flops * B / (kern_ms * 1e-3) / 1e12,
                     "frac_of_fp64_term": (flops * B / (kern_ms * 1e-3) / 1e12) / (fp64_peak * R.world) if fp64_peak else None,
                     "all_correct": ok})
        del full, mine, ct_big
        torch.cuda.empty_cache()
    if not client:
        return None
    return {"metric": "ks_pbs_per_sec", "unit": "keyswitch+PBS/s", "n_gpus": R.world, "value": rows[0]["ks_pbs_per_sec"],
            "workloads": rows,
            "key_broadcast": None if R.world == 1 else
                             {"bytes": sb.key_bytes, "ms": bcast_ms, "gb_per_s": sb.key_bytes / (bcast_ms * 1e-3) / 1e9,
                              "what": "tensor-core keyswitch key (int8 MMA blocks) + multi-bit Fourier bootstrapping key, "
                                      "dist.broadcast over NCCL from the client rank, once per key set"},
            "data_path": "no collective: the client sends rank r its contiguous rows (point-to-point), every rank runs "
                         "fhe_b200_keyswitch_mma + fhe_b200_pbs_mb2 on them, outputs return point-to-point; "
                         "ks_pbs_ms is device time of those two calls (CUDA events, max over ranks), round_trip_ms adds "
                         "the transfers (wall clock between barriers, max over ranks)",
            "server_ranks_hold": "evaluation keys only (no secret key)",
            "all_correct": all(r["all_correct"] for r in rows)}
