#!/usr/bin/env python3
"""bench.py -- encrypted comparisons/s (and PBS/s) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

A *step* is one pass of the encrypted-compare hot path over all resident documents: the encrypted
dot product of every document's ciphertexts with the quantized weights (server), then decrypt +
dequantize of the encrypted scores (client kernels).  `value` times that with ciphertexts already
resident in HBM; `e2e` times the public call with HOST buffers (quantize -> encrypt -> dot -> decrypt on
the device, H2D/D2H inside the timed region) plus the host-side threshold / sort / top-k.

Workloads (--workload auto picks by N):
  N = 1   BASELINE.json configs[1]: `search` top-k=3 over 1000 synthetic encrypted documents (d=128, 8-bit),
          expanded ciphertexts, HBM-bound dot product.  e2e = FHESimilarityModel.predict_encrypted.
          Sub-records: configs[3] on one GPU (1 M seeded documents) and one GPU's share of configs[4].
  N > 1   BASELINE.json configs[3]: ONE collection of 1 M synthetic encrypted documents (seeded ciphertexts),
          contiguous cost-weighted shards over the N GPUs (strong scaling), scores pushed to the client GPU's
          score board.  e2e = ShardedSearch.search(): query in host memory -> ranked ids on the client.
          Sub-records: the 1000-documents-per-GPU weak step of round 1, and configs[4] (d=256, 12-bit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D_FEATURES = 128
N_BITS = 8
# fixed seeds are an explicit opt-in for a reproducible benchmark; the library's defaults come from the OS CSPRNG
DATA_SEED, KEY_SEED, ENC_SEED, EVK_SEED, NOISE_SEED = 20261018, 0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004
# measured DRAM traffic of one lincomb_kernel launch per document, keyed by (lwe n, outputs M, d): dram__bytes_read.sum +
# dram__bytes_write.sum of an `ncu --set full` capture; shapes without a capture report traffic = null
NCU_TRAFFIC = {(1423, 2, 128): (1475440, "profiles/r2_ncu_lincomb_decrypt_v4.txt ((1.458192 GB read + 17.248 MB written) / 1000 documents)")}
TOTAL_DOCS_CONFIG4 = 1_000_000       # BASELINE.json configs[3]
TOTAL_DOCS_CONFIG5 = 100_000         # BASELINE.json configs[4], on 8 GPUs: 12 500 per GPU
D_CONFIG5, BITS_CONFIG5 = 256, 12
DOC_BLOCK = 8192                     # documents per generator block of the large synthetic collections
METRIC = "encrypted_comparisons_per_sec"
UNIT = "comparisons/s"


def build_model(device=None, d=D_FEATURES, n_bits=N_BITS):
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=d, n_bits=n_bits, seed=DATA_SEED, key_seed=KEY_SEED, enc_seed=ENC_SEED,
                           noise_seed=NOISE_SEED, ct_start=0, device=device, verbose=False)
    X, _ = m.train()
    m.compile(X[:10])
    return m, X


def synthetic_docs(n_docs: int, seed: int):
    """Unit-norm query and documents, half of them correlated with the query as in the
    reference's generator (fhe_similarity.py:41-51); X = query * docs (the clear product the
    reference feeds the circuit, batch_operations.py:273)."""
    rng = np.random.RandomState(seed)
    q = rng.randn(D_FEATURES).astype(np.float32)
    q /= np.linalg.norm(q)
    docs = rng.randn(n_docs, D_FEATURES).astype(np.float32)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    mask = rng.rand(n_docs) > 0.5
    docs[mask] = q + 0.2 * rng.randn(int(mask.sum()), D_FEATURES)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    return q, docs, (q[None, :] * docs).astype(np.float32)


def collection_query(d: int, seed: int) -> np.ndarray:
    q = np.random.default_rng(seed).standard_normal(d, dtype=np.float32)
    return q / np.linalg.norm(q)


def collection_rows(lo: int, hi: int, d: int, seed: int, q: np.ndarray) -> np.ndarray:
    """Documents [lo, hi) of a large synthetic collection (unit norm, half of them correlated with the query, as in
    the reference's generator fhe_similarity.py:41-51).  Generated block by block from per-block seeds, so that any
    rank can produce exactly its own rows -- and the client every row, for the exactness check -- whatever the
    sharding."""
    out = np.empty((max(hi - lo, 0), d), dtype=np.float32)
    for blk in range(lo // DOC_BLOCK, (hi + DOC_BLOCK - 1) // DOC_BLOCK if hi > lo else 0):
        rng = np.random.default_rng([seed, blk])
        docs = rng.standard_normal((DOC_BLOCK, d), dtype=np.float32)
        mask = rng.random(DOC_BLOCK) > 0.5
        docs[mask] = q + np.float32(0.2) * rng.standard_normal((int(mask.sum()), d), dtype=np.float32)
        docs /= np.linalg.norm(docs, axis=1, keepdims=True)
        a, b = max(lo, blk * DOC_BLOCK), min(hi, (blk + 1) * DOC_BLOCK)
        out[a - lo: b - lo] = docs[a - blk * DOC_BLOCK: b - blk * DOC_BLOCK]
    return out


def weak_collection_rows(lo: int, hi: int, docs_per_gpu: int) -> np.ndarray:
    """Rows [lo, hi) of the weak-scaling collection: block r = synthetic_docs(docs_per_gpu, DATA_SEED + 1 + r), the
    round-1 per-rank data, concatenated (so that N = 1 is exactly configs[1])."""
    parts = []
    for r in range(lo // docs_per_gpu, (hi + docs_per_gpu - 1) // docs_per_gpu if hi > lo else 0):
        X = synthetic_docs(docs_per_gpu, DATA_SEED + 1 + r)[2]
        a, b = max(lo, r * docs_per_gpu), min(hi, (r + 1) * docs_per_gpu)
        parts.append(X[a - r * docs_per_gpu: b - r * docs_per_gpu])
    return np.concatenate(parts) if parts else np.zeros((0, D_FEATURES), dtype=np.float32)


def top_k(scores: np.ndarray, k: int, min_similarity: float):
    """Reference semantics (batch_operations.py:278-284): filter >=, stable sort descending, [:k]."""
    from fhe_icp_b200.batch_operations import top_indices
    scores = np.asarray(scores, dtype=np.float64)
    return [(int(i), float(scores[i])) for i in top_indices(scores, k, min_similarity)]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------- CPU arm
def cpu_reference(model, X, target_seconds: float, threads=None):
    """Oracle port of predict_encrypted on the host cores: quantize -> encrypt -> dot -> decrypt."""
    from oracle import oracle as O
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its own thread count (all host cores this
    # process may run on) instead of inheriting that
    O.lib().orc_set_num_threads(int(threads) if threads else host_cores())
    c = model.model.fhe_circuit
    spec = c.spec
    s = O.secret_key(c.key_seed, 2, c.lwe.n)
    W = np.stack([spec.q_weights, np.ones_like(spec.q_weights)]) if c.two_outputs else spec.q_weights[None]

    def run(rows):
        reps = (rows + len(X) - 1) // len(X)
        Xr = np.tile(X, (reps, 1))[:rows] if reps > 1 else X[:rows]
        t0 = time.perf_counter()
        q = O.quantize(Xr, spec.input_q.scale, spec.input_q.zero_point, spec.input_q.offset, spec.input_q.n_bits)
        ct = O.lwe_encrypt(s, q, c.lwe.shift, c.lwe.sigma_abs, c.enc_seed, ct_base=0, stride=c.lwe.stride,
                           noise_seed=c.noise_seed)
        out = O.lincomb(ct.reshape(rows, spec.d, -1), W, c.lwe.n)
        m = O.lwe_decrypt(s, out, c.lwe.shift)
        qy = m[:, 0] - (int(spec.weight_q.zero_point) * m[:, 1] if c.two_outputs else 0) + int(spec.q_bias)
        y = spec.dequantize_output(qy)
        return time.perf_counter() - t0, y, Xr

    # bounded sample: chunks of <= 2048 documents (3 GB of host ciphertexts) until ~target_seconds of
    # CPU work have been timed
    chunk = min(2048, max(64, len(X)))
    run(64)  # warm the thread pool / page in the library
    rows, t = 0, 0.0
    while t < target_seconds:
        dt, y, Xr = run(chunk)
        if rows == 0:
            assert np.array_equal(y, model.predict_clear(Xr)), "oracle result != clear circuit"
        rows += chunk
        t += dt
    return rows / t, rows, t, O.num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, _ = build_model()
    wl = resolve_workload(args)
    if wl == "config2":
        X = weak_collection_rows(0, args.docs, args.docs)
        what = f"the {args.docs}-document workload tiled"
    else:   # a bounded sample of the same collection the GPU arm shards: its first documents
        q4 = collection_query(D_FEATURES, DATA_SEED + 77)
        n = min(args.total_docs, 4096)
        X = q4[None, :] * collection_rows(0, n, D_FEATURES, DATA_SEED + 77, q4)
        what = f"the first {n} of the {args.total_docs} documents, tiled"
    vals = []
    per_step = min(args.cpu_seconds, max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup))))
    rows = 0
    threads = 0
    for i in range(args.warmup + args.steps):
        v, rows, _, threads = cpu_reference(model, X, per_step)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    c = model.model.fhe_circuit
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * rows / value, "higher_is_better": True,
        "scaling": "weak" if wl == "config2" else "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, c),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{rows} documents per step ({what}), quantize+encrypt+dot+decrypt, "
                                   "oracle/fhe_oracle.c with OpenMP on all host cores (Concrete itself is not installable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "published_reference": {"value": 5.0, "unit": UNIT, "note": "1/0.20 s per 8-bit sample, hardware unstated "
                                "(reference SESSION_REPORT.md:70)"},
    }
    print(json.dumps(line))
    return 0


def resolve_workload(args) -> str:
    """config2 = BASELINE.json configs[1] (1000 documents per GPU, weak); config4 = configs[3] (one 1 M-document
    collection sharded over the GPUs, strong).  auto: config2 on one GPU, config4 on several."""
    if args.workload != "auto":
        return args.workload
    return "config2" if args.gpus <= 1 else "config4"


def workload_config(args, c):
    """The `config` object of the JSON line.  A pure function of the command line and the compiled circuit, so that
    the GPU arm and the CPU reference arm print the same object for the same N."""
    wl = resolve_workload(args)
    M = 2 if c.two_outputs else 1
    cfg = {"d": D_FEATURES, "n_bits": N_BITS, "lwe_n": c.lwe.n, "ciphertext_words": c.lwe.stride,
           "log2_delta": c.lwe.shift, "log2_sigma": round(c.lwe.log2_sigma, 2), "outputs_per_comparison": M,
           "bytes_per_comparison": comparison_bytes(c), "mask_generator": "Philox4x32-7 (public masks), Philox4x32-10 (secret streams)",
           "seeds": {"data": DATA_SEED, "key": KEY_SEED, "enc": ENC_SEED}}
    if wl == "config2":
        cfg.update({"workload": f"search top-k=3 over {args.docs} synthetic encrypted documents per GPU "
                                "(BASELINE.json configs[1])",
                    "docs_per_gpu": args.docs, "total_docs": args.docs * max(1, args.gpus), "ciphertext_format": "expanded",
                    "l2_policy": "inputs larger than L2 (ciphertext set per step >= 1 GB vs 126 MB L2), no flush"})
    else:
        cfg.update({"workload": f"search top-k=3 over ONE collection of {args.total_docs} synthetic encrypted documents, "
                                f"contiguous document shards over {max(1, args.gpus)} GPU(s) (BASELINE.json configs[3])",
                    "total_docs": args.total_docs, "ciphertext_format": "seeded (8 B per ciphertext + public mask seed)",
                    "l2_policy": "no reuse between steps: masks are regenerated, never read; bodies 1 KB per document"})
    if args.gpus > 1:
        cfg["multi_gpu"] = ("one process per GPU; every rank evaluates its shard and its dot-product kernel stores the encrypted "
                            "scores (32-bit wire form) into the client GPU's score board over NVLink (cudaIpc peer memory) and "
                            "flags their arrival; the client decrypts them under the next step's dot products; no data-path "
                            "collective (NCCL: setup and the max-over-ranks timing only)")
    return cfg


def comparison_bytes(c) -> int:
    M = 2 if c.two_outputs else 1
    return (c.spec.d + M) * (c.lwe.n + 1) * 8


# ------------------------------------------------------------------------------------- GPU arm
class Ranks:
    """Process-group facts of this run (one process per GPU)."""

    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dev = torch.device("cuda", self.local_rank)

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def measure_client_rho(model, fmt: str, sample_docs: int = 4000) -> float:
    """Client cost per document / server cost per document, measured on this GPU: the fused decrypt of the 32-bit wire
    form against the dot product of the same documents.  Feeds sharded_search.client_cost_weights: the client rank
    decrypts every document of the collection, so it takes a smaller shard."""
    import torch
    c = model.model.fhe_circuit
    M = 2 if c.two_outputs else 1
    dev = model.dev
    X = weak_collection_rows(0, sample_docs, sample_docs) if c.spec.d == D_FEATURES else \
        np.zeros((sample_docs, c.spec.d), dtype=np.float32)
    saved = c.ct_counter
    c.ct_counter = 1 << 50
    ct = model.encrypt(X, seeded=(fmt == "seeded"))
    c.ct_counter = saved
    out = torch.empty((sample_docs, M, c.lwe.stride), dtype=torch.int64, device=dev)
    out32 = torch.empty((sample_docs, M, c.lwe.stride), dtype=torch.int32, device=dev)

    def ev(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_srv = ev(lambda: model.run(ct, out=out))
    model.compress_scores(out, out32)
    t_cli = ev(lambda: _decrypt_device(model, out32, wire32=True))
    return t_cli / t_srv


def sharded_steps(R: Ranks, model, rows_fn, total_docs: int, fmt: str, steps: int, warmup: int, gather_mode: str,
                  rho: float, ctx, verify_chunk: int = 65536):
    """The device-resident measurement: `total_docs` documents in contiguous (cost-weighted) shards, every rank's
    ciphertexts resident; K timed steps of dot products (+ gather to the client + client decrypt).  Returns a dict;
    `value` = total_docs * K / max-over-ranks time.

    rows_fn(lo, hi) -> float32 [hi-lo, d]: the clear products X of documents [lo, hi) (what the reference feeds the
    circuit, batch_operations.py:273)."""
    import torch
    import torch.distributed as dist
    from fhe_icp_b200.sharded_search import client_cost_weights, shard_bounds
    c = model.model.fhe_circuit
    d, M = c.spec.d, (2 if c.two_outputs else 1)
    rank, world, dev = R.rank, R.world, R.dev
    weights = client_cost_weights(world, 0, rho)
    spans = [shard_bounds(total_docs, world, r, weights) for r in range(world)]
    lo, hi = spans[rank]
    B = hi - lo
    rows_max = max(h - l for l, h in spans)
    X = rows_fn(lo, hi)
    seeded = fmt == "seeded"
    c.ct_counter = lo * d                 # ciphertext (document, j) has id document*d + j whatever the sharding
    ct = model.encrypt(X, seeded=seeded) if B else None
    torch.cuda.synchronize()

    mode = gather_mode if world > 1 else "local"
    if mode == "push" and not c.wire32_supported:
        mode = "gather"                   # the score board carries the 32-bit wire form only
    wire32 = bool(c.wire32_supported)
    board = None
    if mode == "push":
        from fhe_icp_b200.score_board import PeerScoreBoard
        try:   # setup is collective and fails on every rank together (score_board.py), so all ranks take the same path
            board = PeerScoreBoard(model, rows_max, client_rank=0)
        except Exception as e:  # e.g. no peer access between the GPUs of this box: NCCL path instead
            print(f"bench.py: rank {rank}: peer score board unavailable ({e}); using all_gather", file=sys.stderr)
            board, mode = None, "all_gather"
    nccl = mode in ("gather", "all_gather")
    nbuf = 2 if nccl else 1
    # NCCL forms need equal contributions: outputs padded to the largest shard
    outs = [torch.zeros((rows_max if nccl else max(B, 1), M, c.lwe.stride), dtype=torch.int64, device=dev) for _ in range(nbuf)] \
        if mode != "push" else []
    wdt = torch.int32 if wire32 else torch.int64
    outs32 = [torch.zeros(o.shape, dtype=torch.int32, device=dev) for o in outs] if (nccl and wire32) else None
    allg = mode == "all_gather"
    gathered = [torch.empty((world * rows_max, M, c.lwe.stride), dtype=wdt, device=dev) for _ in outs] \
        if (nccl and (rank == 0 or allg)) else None
    post = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    done = [None, None]
    last = [None]

    def client_decrypt(buf):
        for r, (l, h) in enumerate(spans):
            if h > l:
                _decrypt_device(model, buf[r * rows_max: r * rows_max + (h - l)], wire32=wire32, slot=r)

    def step(i, ev=None):
        cur = torch.cuda.current_stream(dev)
        if mode == "push":
            if ev:
                ev[0].record()
            board.push(ct)                               # server: dot products, scores pushed to the client
            if ev:
                ev[1].record()
            if rank == 0:
                with torch.cuda.stream(post):
                    last[0] = board.collect()            # client: wait for every shard's arrival flag,
                    client_decrypt(last[0])              # decrypt every shard's scores,
                    board.release()                      # hand the slot back
            return
        k = i % len(outs)
        if world > 1 and done[k] is not None:
            cur.wait_event(done[k])                      # buffer k was consumed by the post stream
        if ev:
            ev[0].record()
        if B:
            model.run(ct, out=outs[k][:B])               # server: encrypted dot products of this shard
        if ev:
            ev[1].record()
        if world == 1:
            _decrypt_device(model, outs[k][:B])          # client kernels: decrypt + dequantize
            return
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(post):
            post.wait_event(ready)
            send = outs[k]
            if wire32:
                model.compress_scores(outs[k], outs32[k])
                send = outs32[k]
            if allg:
                dist.all_gather_into_tensor(gathered[k], send)
            else:
                dist.gather(send, list(gathered[k].chunk(world)) if rank == 0 else None, dst=0)
            if rank == 0:
                last[0] = gathered[k]
                client_decrypt(gathered[k])              # client: decrypt every shard's scores
            done[k] = torch.cuda.Event()
            done[k].record(post)

    for i in range(max(warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    R.barrier()
    # exactness: every shard's scores, as the client decrypted them, equal the clear quantized circuit
    if mode == "push":
        board.check()
    if rank == 0:
        for r, (l, h) in enumerate(spans):
            for a in range(l, h, verify_chunk):
                b = min(h, a + verify_chunk)
                ref = model.predict_clear(X[a - lo: b - lo] if r == rank else rows_fn(a, b))
                if world == 1:
                    got = model.decrypt(outs[0][a - l: b - l])
                else:
                    enc = last[0][r * rows_max + (a - l): r * rows_max + (b - l)]
                    got = model.decrypt_compressed(enc) if wire32 else model.decrypt(enc)
                assert np.array_equal(got, ref), f"shard {r}, documents [{a},{b}): GPU scores differ from the clear quantized circuit"
    R.barrier()

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sampler = ClockSampler(R.local_rank)
    if rank == 0:
        sampler.start()
    R.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        step(i, evs[i])
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(post)
    t1.record()
    torch.cuda.synchronize()
    R.barrier()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = R.max_over_ranks(t0.elapsed_time(t1))
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if board is not None:
        board.check()      # no in-stream wait timed out
        board.close()
    return {"value": total_docs * steps / (elapsed_ms * 1e-3), "elapsed_ms": elapsed_ms, "ms_per_step": elapsed_ms / steps,
            "kern_ms": kern_ms, "kern_docs": B, "launches": int(launches), "mode": mode, "clocks": clocks,
            "shards": [h - l for l, h in spans], "client_rho": rho, "X": X, "wire32": wire32}


def e2e_single_gpu(model, X, steps, blocks=5):
    """FHESimilarityModel.predict_encrypted with host float32 rows in / float64 scores out + host top-k.
    The call takes ~0.36 ms, so ONE host hiccup (a descheduled thread, a page fault) inside a 10-call loop moves the figure
    by tens of percent (2.0 vs 2.8 M/s were both seen on the pool's boxes).  The loop is therefore timed in `blocks`
    blocks of `steps` calls and the MEDIAN block is reported; every block's figure is kept in the record."""
    import torch
    from fhe_icp_b200._native import pinned_copy
    c = model.model.fhe_circuit
    res, per_block = {}, {}
    X = pinned_copy(np.ascontiguousarray(X, dtype=np.float32))    # the step's inputs sit in pinned host memory
    for fmt in ("expanded", "seeded"):        # seeded (the default format) last: its ranking is the one checked
        c.ciphertext_format = fmt
        for _ in range(3):
            model.predict_encrypted(X)
        vals = []
        for _ in range(blocks):
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for _ in range(steps):
                sc = model.predict_encrypted(X)
                hits = top_k(sc, 3, -np.inf)
            torch.cuda.synchronize()
            vals.append(len(X) * steps / (time.perf_counter() - w0))
        res[fmt] = float(np.median(vals))
        per_block[fmt] = [float(v) for v in vals]
    res["blocks"] = per_block
    return res, hits


def e2e_sharded(R: Ranks, model, total_docs, rho, steps, seed):
    """ShardedSearch.search(): the query and the collection are in HOST memory (pinned); per query every rank uploads
    its shard's rows, the encryption kernel forms the clear products query * doc, quantizes and encrypts them, the rank
    evaluates and pushes; the client decrypts, reads the scores back and ranks.  All GPUs belong to the key owner
    (key_holders='all', the reference's one-machine trust model), so the client-side encryption is sharded too.
    A second figure keeps the (plaintext, client-owned) collection resident on the GPUs: only the query is uploaded."""
    import torch
    from fhe_icp_b200.sharded_search import ShardedSearch, client_cost_weights, shard_bounds
    c = model.model.fhe_circuit
    d = c.spec.d
    weights = client_cost_weights(R.world, 0, rho)
    lo, hi = shard_bounds(total_docs, R.world, R.rank, weights)
    q = collection_query(d, seed)
    docs = collection_rows(lo, hi, d, seed, q)
    c.ct_counter = (1 << 44) + R.rank * (1 << 40)       # fresh ids, disjoint between the ranks
    c.ciphertext_format = "seeded"
    gather = "push" if (R.world > 1 and c.wire32_supported) else "nccl"
    vals = {}
    for placement in ("device", "host"):        # "host" (everything uploaded per query) last: it is the headline
        ss = ShardedSearch(model, docs, key_holders="all", gather=gather, shard_weights=weights, n_docs=total_docs,
                           collection=placement)
        res = None
        for _ in range(2):
            res = ss.search(q, top_k=3, min_similarity=-np.inf)
        R.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(steps):
            res = ss.search(q, top_k=3, min_similarity=-np.inf)
        torch.cuda.synchronize()
        vals[placement] = total_docs * steps / R.max_over_ranks(time.perf_counter() - w0)
        ss.close()
        del ss
    ok = None
    if R.rank == 0:     # the ranking equals the clear circuit's over the whole collection
        ref = np.concatenate([model.predict_clear(q[None, :] * collection_rows(a, min(total_docs, a + 65536), d, seed, q))
                              for a in range(0, total_docs, 65536)])
        ok = [i for i, _ in top_k(ref, 3, -np.inf)] == [int(name.split("_")[1]) for name, _ in res]
        assert ok, "sharded search ranking differs from the clear circuit"
    return {"value": vals["host"], "unit": UNIT, "steps": steps,
            "h2d_bytes_per_step": int(total_docs * d * 4 + d * 4), "d2h_bytes_per_step": int(total_docs * 8),
            "call": f"ShardedSearch(key_holders='all', gather='{gather}', cost-weighted shards).search(query, top_k=3): query "
                    "and documents in pinned host memory -> H2D of every shard's rows -> product + quantize + encrypt "
                    "(seeded) -> dot products -> scores pushed to the client GPU -> client decrypt -> D2H -> host top-k",
            "resident_collection_value": vals["device"],
            "resident_collection_note": "same call with collection='device': the client-owned plaintext rows stay on the "
                                        f"GPUs, {d * 4} B of query go up per step",
            "ranking_equals_clear_circuit": ok}


def hbm_roofline(c, kern_ms, kern_docs, peak, peak_src):
    M = 2 if c.two_outputs else 1
    bytes_per_launch = comparison_bytes(c) * kern_docs
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    t = NCU_TRAFFIC.get((c.lwe.n, M, c.spec.d))
    return {"bound": "hbm", "kernel": "lincomb_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": t[0] * kern_docs if t else None,
            "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, " + t[1]) if t else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_per_launch), "kernel_ms": kern_ms,
            "documents_per_launch": int(kern_docs)}


def seeded_roofline(c, kern_ms, kern_docs, sm_mhz):
    """The seeded dot product reads 1 KB per document: no HBM roofline applies.  It is bound by the wide integer
    multiplier (IMAD.WIDE.U32 on the fmaheavy pipe): 12 per Philox block of two mask words (10 of Philox4x32-7's 14
    -- rounds 1-2 are shared per block and per ciphertext -- plus 2 for the weighted sums)."""
    d, n = c.spec.d, c.lwe.n
    blocks = kern_docs * d * ((n + 2) // 2)
    wide = 12.0 * blocks
    achieved = wide / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    mhz = sm_mhz or 1965.0
    peak = 148 * 4 * 32 / 4.0 * mhz * 1e6 / 1e12        # one IMAD.WIDE per 4 cycles per scheduler, 32 lanes
    return {"bound": "integer (wide-multiplier pipe, fmaheavy)", "kernel": "lincomb_seeded_kernel", "achieved": achieved,
            "peak": peak, "unit": "T wide-multiplies/s", "frac": achieved / peak, "traffic": None,
            "peak_source": "issue rate of IMAD.WIDE.U32 taken as 1 per 4 cycles per scheduler (16-lane fmaheavy pipe) at the "
                           "sampled SM clock; ncu: sm__pipe_fmaheavy_cycles_active 86 % (profiles/r2_ncu_e2e_seeded_v1.txt), "
                           "sustained rate inside Philox rounds 1 per 7 cycles (profiles/r2_seeded_kernel_times.txt)",
            "kernel_ms": kern_ms, "documents_per_launch": int(kern_docs), "philox_blocks_per_launch": int(blocks),
            "docs_per_sec_per_gpu": kern_docs / (kern_ms * 1e-3) if kern_ms > 0 else None,
            "equivalent_expanded_stream_gbs": comparison_bytes(c) * kern_docs / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from fhe_icp_b200 import _native as N

    R = Ranks()
    rank, world, dev = R.rank, R.world, R.dev
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(R.local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the contract is ONE JSON line on stdout, and NCCL logs to stdout by default: its log (INFO unless the caller
        # chose a level) goes to stderr instead, where a driver can still read the communicator / rank lines
        # (the GPU boxes export NCCL_DEBUG=VERSION, whose banner would land on stdout: anything below INFO is raised).
        # Every rank logs into its own file and copies it to stderr when it is done: N processes that each open
        # /dev/stderr themselves overwrite one another when stderr is a regular file.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")     # communicator / rank / transport lines only
        if "NCCL_DEBUG_FILE" not in os.environ:
            import atexit
            import tempfile
            nccl_log = os.path.join(tempfile.gettempdir(), f"bench_nccl_{os.getpid()}.log")
            os.environ["NCCL_DEBUG_FILE"] = nccl_log

            def _dump_nccl_log(path=nccl_log):
                try:
                    with open(path, errors="replace") as f:
                        sys.stderr.write(f.read())
                    sys.stderr.flush()
                    os.unlink(path)
                except OSError:
                    pass
            atexit.register(_dump_nccl_log)
        # NCCL's stream (and the post stream) run at high priority: the dot-product kernel keeps
        # thousands of CTAs queued, and equal-priority kernels only start once those are all dispatched
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    args.gpus = world
    wl = resolve_workload(args)
    model, _ = build_model(device=R.local_rank)
    c = model.model.fhe_circuit
    ctx = N.context(R.local_rank)
    peak, peak_src = measured_peak_gbs()
    warm = max(args.warmup, 3)

    rho_cache = {}

    def rho_for(fmt):
        if world == 1 or args.client_rho == 0:
            return 0.0
        if args.client_rho > 0:
            return args.client_rho
        # measured on the client GPU, padded for what the measurement cannot see (inbound NVLink stores landing in the
        # client's HBM while it streams its own shard, the decrypt CTAs' SM slots), same value on every rank
        if fmt not in rho_cache:
            r = measure_client_rho(model, fmt) * args.client_rho_pad if rank == 0 else 0.0
            rho_cache[fmt] = R.max_over_ranks(r)
        return rho_cache[fmt]

    sub = {}
    if wl == "config2":
        fmt = args.format
        total = args.docs * world
        res = sharded_steps(R, model, lambda lo, hi: weak_collection_rows(lo, hi, args.docs), total, fmt, args.steps, warm,
                            args.gather_mode, rho_for(fmt), ctx)
        scaling = "weak"
    else:
        fmt = "seeded"
        total = args.total_docs
        q4 = collection_query(D_FEATURES, DATA_SEED + 77)
        rows4 = lambda lo, hi: q4[None, :] * collection_rows(lo, hi, D_FEATURES, DATA_SEED + 77, q4)   # noqa: E731
        res = sharded_steps(R, model, rows4, total, fmt, args.steps, warm, args.gather_mode, rho_for(fmt), ctx)
        scaling = "strong"
    X = res.pop("X")

    # --- e2e: host buffers in, ranked ids out
    e2e_steps = max(3, min(args.steps, 20))
    if world == 1 and wl == "config2":
        e2e_vals, hits = e2e_single_gpu(model, X, e2e_steps)
        assert [i for i, _ in hits] == [i for i, _ in top_k(model.predict_clear(X), 3, -np.inf)], \
            "top-k ranking differs from the clear circuit"
        e2e = {"value": e2e_vals["seeded"], "unit": UNIT, "h2d_bytes_per_step": int(len(X) * D_FEATURES * 4),
               "d2h_bytes_per_step": int(len(X) * 16), "steps": e2e_steps,
               "timing": f"median of {len(e2e_vals['blocks']['seeded'])} timed blocks of {e2e_steps} calls each (wall clock, "
                         "synchronised on both sides of a block)",
               "blocks": e2e_vals["blocks"]["seeded"],
               "call": "fhe_b200_similarity_predict_host_seeded (FHESimilarityModel.predict_encrypted, default "
                       "ciphertext_format='seeded') + host top-k",
               "expanded_ciphertexts_value": e2e_vals["expanded"]}
    else:
        del X
        e2e = e2e_sharded(R, model, total if wl == "config4" else args.docs * world, rho_for("seeded"),
                          max(3, min(e2e_steps, 5)), DATA_SEED + 77)

    # --- sub-records: the other configurations BASELINE.json names, each with its own exactness check
    if not args.no_sub_records:
        torch.cuda.empty_cache()
        if wl == "config4":      # the round-1 weak step (1000 expanded documents per GPU), for the scaling history
            r2 = sharded_steps(R, model, lambda lo, hi: weak_collection_rows(lo, hi, args.docs), args.docs * world, "expanded",
                               args.steps, warm, args.gather_mode, rho_for("expanded"), ctx)
            r2.pop("X")
            sub["config2_weak_1000_docs_per_gpu"] = dict(
                {k: r2[k] for k in ("value", "ms_per_step", "mode", "shards", "client_rho")}, unit=UNIT, scaling="weak",
                docs_per_gpu=args.docs, roofline=hbm_roofline(c, r2["kern_ms"], r2["kern_docs"], peak, peak_src))
        elif args.total_docs > 0:   # configs[3] on this one GPU: the N = 1 point of the strong-scaling curve
            q4 = collection_query(D_FEATURES, DATA_SEED + 77)
            rows4 = lambda lo, hi: q4[None, :] * collection_rows(lo, hi, D_FEATURES, DATA_SEED + 77, q4)   # noqa: E731
            r4 = sharded_steps(R, model, rows4, args.total_docs, "seeded", max(3, min(args.steps, 5)), 3, "local", 0.0, ctx)
            r4.pop("X")
            sub["config4_1M_docs_seeded_one_gpu"] = dict(
                {k: r4[k] for k in ("value", "ms_per_step")}, unit=UNIT, total_docs=args.total_docs,
                roofline=seeded_roofline(c, r4["kern_ms"], r4["kern_docs"], (res["clocks"] or {}).get("sm_mhz")))
        torch.cuda.empty_cache()
        docs5 = args.docs5_per_gpu
        if docs5 > 0:            # configs[4]: d = 256, 12-bit; 100 k documents over 8 GPUs = 12 500 per GPU
            m5, _ = build_model(device=R.local_rank, d=D_CONFIG5, n_bits=BITS_CONFIG5)
            c5 = m5.model.fhe_circuit
            q5 = collection_query(D_CONFIG5, DATA_SEED + 78)
            rows5 = lambda lo, hi: q5[None, :] * collection_rows(lo, hi, D_CONFIG5, DATA_SEED + 78, q5)   # noqa: E731
            r5 = sharded_steps(R, m5, rows5, docs5 * world, "expanded", max(3, min(args.steps, 10)), 3, args.gather_mode,
                               0.0, ctx, verify_chunk=4096)
            r5.pop("X")
            sub["config5_d256_12bit"] = dict(
                {k: r5[k] for k in ("value", "ms_per_step", "mode", "shards", "wire32")}, unit=UNIT,
                workload=f"search over {docs5 * world} synthetic encrypted documents, d={D_CONFIG5}, n_bits={BITS_CONFIG5}, "
                         f"{docs5} per GPU ({TOTAL_DOCS_CONFIG5} on 8 GPUs is BASELINE.json configs[4])",
                lwe_n=c5.lwe.n, log2_delta=c5.lwe.shift, bytes_per_comparison=comparison_bytes(c5),
                exact_vs_clear_circuit=True, roofline=hbm_roofline(c5, r5["kern_ms"], r5["kern_docs"], peak, peak_src))
            del m5, c5
            torch.cuda.empty_cache()

    # --- keyswitch + PBS over one batch sharded across the ranks (the metric's "PBS/sec at 1/2/4/8 B200"): every rank takes part
    pbs_sharded = None
    if not args.no_extras:
        from fhe_icp_b200 import pbs_bench as _pbs_bench
        torch.cuda.empty_cache()
        pbs_sharded = _pbs_bench.measure_sharded(R, dev, args)
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    args._gather_mode = res["mode"]
    clocks = res["clocks"]
    roof = seeded_roofline(c, res["kern_ms"], res["kern_docs"], (clocks or {}).get("sm_mhz")) if fmt == "seeded" \
        else hbm_roofline(c, res["kern_ms"], res["kern_docs"], peak, peak_src)
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(args, c),
        "e2e": e2e, "gpu_launches": res["launches"], "roofline": roof, "clocks": clocks,
        "gather_mode": res["mode"], "shards": res["shards"], "client_rho": res["client_rho"],
        "exact_vs_clear_circuit": True, "sub_records": sub,
    }
    if not args.no_cpu_baseline and world == 1:
        Xc = weak_collection_rows(0, args.docs, args.docs)
        v, rows, t, threads = cpu_reference(model, Xc, args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{rows} documents (the {args.docs}-document workload tiled), quantize+encrypt+dot+decrypt in {t:.1f} s, "
                                          "oracle/fhe_oracle.c (OpenMP)"}
    try:
        if args.no_extras:
            raise ImportError
        from fhe_icp_b200 import pbs_bench
        args._hbm_peak = peak
        try:   # dense bf16 burst rate measured on this pool (denominator of the tensor-core keyswitch roofline)
            args._bf16_peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
        except Exception:
            args._bf16_peak = None
        line["pbs"] = pbs_bench.measure(dev, args)
        line["pbs_sharded"] = pbs_sharded
        if world == 1 and not args.no_cpu_baseline:
            line["pbs"]["cpu_baseline"] = cpu_pbs_reference(line["pbs"]["params"])
        if world == 1:
            pair = pbs_bench.measure_pair(dev, args)
            sample = pair.pop("_sample")
            if not args.no_cpu_baseline:
                pair["cpu_baseline"] = cpu_pair_reference(pair["params"], sample)
            line["encrypted_pair"] = pair
            packed = pbs_bench.measure_packed(dev, args)
            sample = packed.pop("_sample")
            if not args.no_cpu_baseline:
                packed["cpu_baseline"] = cpu_packed_reference(packed["params"], sample)
            line["encrypted_pair_packed"] = packed
    except ImportError:
        pass
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_pbs_reference(params, per_thread: int = 16):
    """CPU baseline of configs[2] (keyswitch + PBS at the stated set): the oracle port, OpenMP over the host cores, on a
    bounded batch (``per_thread`` ciphertexts per thread), multi-bit blind rotation like the GPU path; every output is
    decrypted and compared with the table."""
    import numpy as np
    from oracle import oracle as O
    op = O.make_params(n=params["n"], k=params["k"], N=params["N_poly"], l_pbs=params["l_pbs"],
                       beta_pbs=params["beta_pbs"], l_ks=params["l_ks"], beta_ks=params["beta_ks"],
                       log2_sigma_lwe=params["log2_sigma_lwe"], log2_sigma_glwe=params["log2_sigma_glwe"])
    s, S = O.secret_key(101, 0, op.n), O.secret_key(101, 1, op.k * op.N)
    ksk = O.ksk_gen(op, S, s, 202)
    bskf2 = O.bsk2_to_fourier(op, O.bsk2_gen(op, s, S, 202))
    B = per_thread * O.num_threads()
    msgs = np.random.RandomState(B).randint(0, 16, size=B)
    table = (np.arange(16) * 7 + 3) % 16
    lut = O.make_lut_poly(table, 4, op.N, 59)
    ct = O.lwe_encrypt(S, msgs, 59, op.sigma_glwe_abs, 303, ct_base=0, stride=op.N + 2)[:, : op.N + 1]
    O.pbs_mb2(op, bskf2, O.keyswitch(op, ksk, ct[: O.num_threads()]), lut)      # warm-up (thread pool, key pages)
    t0 = time.perf_counter()
    small = O.keyswitch(op, ksk, ct)
    t1 = time.perf_counter()
    out = O.pbs_mb2(op, bskf2, small, lut)
    t2 = time.perf_counter()
    pad = np.zeros((B, op.N + 2), dtype=np.uint64)
    pad[:, : op.N + 1] = out
    ok = bool(np.array_equal(O.lwe_decrypt(S, pad, 59) & 15, table[msgs]))
    return {"value": B / (t2 - t0), "unit": "keyswitch+PBS/s", "pbs_per_sec": B / (t2 - t1), "cores": O.num_threads(),
            "kind": "port", "all_correct": ok,
            "sample": f"{B} ciphertexts: keyswitch {t1 - t0:.2f} s + multi-bit PBS {t2 - t1:.2f} s, oracle/fhe_oracle.c (OpenMP; a "
                      "plain radix-2 f64 FFT restatement, not a tuned CPU library)"}


def cpu_pair_reference(params, sm):
    """CPU baseline of the encrypted x encrypted comparison: the oracle port on two documents
    (d = 128 programmable bootstraps each with the squared-norm protocol, OpenMP over the host cores)."""
    import numpy as np
    from oracle import oracle as O
    op = O.make_params(n=params["n"], k=params["k"], N=params["N_poly"], l_pbs=params["l_pbs"],
                       beta_pbs=params["beta_pbs"], l_ks=params["l_ks"], beta_ks=params["beta_ks"],
                       log2_sigma_lwe=params["log2_sigma_lwe"], log2_sigma_glwe=params["log2_sigma_glwe"])
    xq, yq = sm["xq"], sm["yq"]
    nd, d = yq.shape
    os_, oS = O.secret_key(sm["key_seed"], 0, op.n), O.secret_key(sm["key_seed"], 1, op.k * op.N)
    obskf = O.bsk2_to_fourier(op, O.bsk2_gen(op, os_, oS, sm["evk_seed"]))
    oq = O.lwe_encrypt(os_, xq, sm["in_shift"], op.sigma_lwe_abs, 1, 0, stride=sm["stride"])
    od = O.lwe_encrypt(os_, yq, sm["in_shift"], op.sigma_lwe_abs, 1, d, stride=sm["stride"]).reshape(nd, d, -1)
    onq = O.lwe_encrypt(oS, (xq * xq).sum(), sm["out_shift"] - 1, op.sigma_glwe_abs, 1, 1 << 40)
    ond = O.lwe_encrypt(oS, (yq * yq).sum(axis=1), sm["out_shift"] - 1, op.sigma_glwe_abs, 1, (1 << 40) + 1)
    t0 = time.perf_counter()
    ref = O.encrypted_product_scores_norms(op, obskf, oq, od, onq, ond, sm["p_bits"], sm["out_shift"], multibit=True)
    cpu_s = time.perf_counter() - t0
    dec = O.lwe_decrypt(oS, ref, sm["out_shift"]) & 8191
    dec = np.where(dec >= 4096, dec - 8192, dec)
    return {"value": nd / cpu_s, "unit": "comparisons/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{nd} documents ({nd * d} PBS) in {cpu_s:.1f} s, oracle/fhe_oracle.c (OpenMP)",
            "agrees_with_gpu": bool(np.array_equal(dec, np.asarray(sm["expect"])))}


def cpu_packed_reference(params, sm):
    """CPU baseline of the packed both-encrypted comparison: the oracle's external product (OpenMP over the
    host cores) on 64 ciphertexts = 1024 documents."""
    import numpy as np
    from oracle import oracle as O
    op = O.make_params(n=params["n"], k=params["k"], N=params["N_poly"], l_pbs=params["l_pbs"],
                       beta_pbs=params["beta_pbs"], l_ks=params["l_ks"], beta_ks=params["beta_ks"],
                       log2_sigma_lwe=params["log2_sigma_lwe"], log2_sigma_glwe=params["log2_sigma_glwe"])
    oS = O.secret_key(sm["key_seed"], 1, op.k * op.N)
    docs = O.glwe_encrypt_rows(op, oS, O.pack_documents(sm["yq"], op.N, sm["slot"]), 0, sm["out_shift"], 1, 1 << 20)
    gf = O.ggsw_to_fourier(op, O.glwe_encrypt_rows(op, oS, O.query_polynomial(sm["xq"], op.N), 1, 0, 2, 0))
    O.glwe_external_product(op, gf, docs[:2])
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        prod = O.glwe_external_product(op, gf, docs)
    cpu_s = (time.perf_counter() - t0) / reps
    lwe = O.glwe_sample_extract(op, prod, 0, sm["slot"], sm["per"], op.N + 2)
    dec = O.lwe_decrypt(oS, lwe, sm["out_shift"]) & 131071
    dec = np.where(dec >= 65536, dec - 131072, dec)
    n = len(sm["yq"])
    return {"value": n / cpu_s, "unit": "comparisons/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{n} documents ({docs.shape[0]} external products) in {cpu_s * 1e3:.1f} ms, oracle/fhe_oracle.c (OpenMP)",
            "agrees_with_gpu": bool(np.array_equal(dec[:n], np.asarray(sm["expect"])))}


def _decrypt_device(model, out, wire32=False, slot=0):
    """decrypt + dequantize kernels without the device->host copy (device-resident timing)."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N
    c = model.model.fhe_circuit
    B = out.shape[0]
    bufs = model.__dict__.setdefault("_bench_bufs", {})
    if slot not in bufs or bufs[slot][0].shape[0] < B:
        bufs[slot] = (torch.empty(B, dtype=torch.float64, device=out.device), torch.empty(B, dtype=torch.int64, device=out.device))
    y, qy = bufs[slot]
    st = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
    fn = N.lib().fhe_b200_similarity_decrypt32 if wire32 else N.lib().fhe_b200_similarity_decrypt
    N.check(fn(c.handle, C.c_void_p(out.data_ptr()), B, C.c_void_p(y.data_ptr()), C.c_void_p(qy.data_ptr()), st))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=1000, help="documents per GPU of the configs[1] (weak) workload")
    ap.add_argument("--workload", default="auto", choices=["auto", "config2", "config4"],
                    help="auto: BASELINE.json configs[1] on one GPU, configs[3] (one 1 M-document collection, sharded) on several")
    ap.add_argument("--total-docs", type=int, default=TOTAL_DOCS_CONFIG4, help="documents of the configs[3] collection")
    ap.add_argument("--docs5-per-gpu", type=int, default=TOTAL_DOCS_CONFIG5 // 8,
                    help="documents per GPU of the configs[4] sub-record (d=256, 12-bit); 0 skips it")
    ap.add_argument("--client-rho", type=float, default=-1.0,
                    help="client cost per document / server cost per document for the cost-weighted shards "
                         "(sharded_search.client_cost_weights); < 0: measured on the client GPU, 0: equal shards")
    ap.add_argument("--client-rho-pad", type=float, default=1.5,
                    help="factor on the measured client/server cost ratio (inbound NVLink stores and decrypt CTAs compete "
                         "with the client's own shard)")
    ap.add_argument("--no-sub-records", action="store_true", help="only the headline workload")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--format", default="expanded", choices=["expanded", "seeded"],
                    help="resident ciphertext format for the device-timed step: expanded = full (n+1)-word ciphertexts "
                         "(HBM-bound dot product, the headline); seeded = 8-byte bodies, masks regenerated on the fly "
                         "(integer-bound; what makes the 1M-document configuration fit)")
    ap.add_argument("--gather-mode", default="push", choices=["push", "gather", "all_gather"],
                    help="how encrypted scores reach the client rank (N>1): push = stored by the dot-product kernel "
                         "into the client's memory over NVLink (peer score board); gather / all_gather = NCCL")
    ap.add_argument("--pbs-batch", type=int, default=0, help="PBS microbench batch (0 = default sweep)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the PBS / encrypted-pair sections that follow the timed loop (for a clean ncu launch list)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
