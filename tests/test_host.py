"""Host-side logic and the C-ABI surface (no GPU): the FHESimilarityModel / estimator mirror of
the reference interface, quantizers, parameter selection, and that libfhe_b200.so loads and
exports every symbol include/fhe_b200.h declares."""
import ctypes as C
import pickle
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


# ------------------------------------------------------------------------------- C-ABI
def test_library_exports_every_declared_symbol():
    from fhe_icp_b200 import _native as N
    header = (ROOT / "include" / "fhe_b200.h").read_text()
    declared = set(re.findall(r"\b(fhe_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = N.lib()  # builds if stale, binds argtypes for every entry of N.SIGNATURES
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fhe_b200.h but not exported"
    assert declared == set(N.SIGNATURES), "ctypes table and header disagree"
    assert lib.fhe_b200_abi_version() == 2


def test_struct_layouts_match_header():
    from fhe_icp_b200 import _native as N
    assert C.sizeof(N.PBSParams) == 48          # 8 x int32 + 2 x double
    assert C.sizeof(N.SimilaritySpec) == 6 * 4 + 2 * 8 + 4 * 8 + 8 + 8 + 8 + 8
    assert N.SimilaritySpec.key_seed.offset == 88 and N.SimilaritySpec.noise_seed.offset == 96


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fhe_icp_b200 import _native as N
    with pytest.raises(N.FheB200Error) as e:
        N.Context(0)
    assert e.value.code == N.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=8, n_bits=8, seed=0, verbose=False)
    X, _ = m.train(n_samples=50)
    m.compile(X[:10])
    with pytest.raises(N.FheB200Error):
        m.predict_encrypted(X[:2])        # fhe="execute" must fail loudly, never fall back
    assert m.predict_clear(X[:2]).shape == (2,)


def test_product_never_imports_the_oracle():
    for f in (ROOT / "fhe_icp_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = f.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
            assert not re.search(r"#\s*include\s*[<\"].*oracle", txt), f
            assert "liboracle" not in txt, f


def test_invalid_arguments_return_errors_not_crashes():
    from fhe_icp_b200 import _native as N
    lib = N.lib()
    assert lib.fhe_b200_ctx_create(0, None) == N.ERR_INVALID
    assert lib.fhe_b200_lincomb(None, None, 1, 1, 1, 2, None, 1, None, 0, None, None) == N.ERR_INVALID
    assert b"null ctx" in lib.fhe_b200_last_error()
    assert lib.fhe_b200_ksk_words(None) == 0


# ------------------------------------------------------------------------------- model surface
def test_model_surface_metrics_and_errors(tmp_path):
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=42, verbose=False)
    with pytest.raises(RuntimeError, match="Model not trained"):
        m.compile(np.zeros((2, 128), dtype=np.float32))
    with pytest.raises(RuntimeError, match="Model not trained"):
        m.predict_clear(np.zeros((2, 128), dtype=np.float32))
    with pytest.raises(RuntimeError, match="Model not compiled"):
        m.predict_encrypted(np.zeros((2, 128), dtype=np.float32))
    X, y = m.train()
    assert X.shape == (1000, 128) and X.dtype == np.float32 and y.shape == (1000,)
    assert {"train_time", "train_score"} <= set(m.metrics) and m.metrics["train_score"] > 0.99
    m.compile(X[:10])
    assert m.compiled and {"compile_time", "compile_memory_mb", "circuit_max_bits"} <= set(m.metrics)
    assert m.model.fhe_circuit.graph.maximum_integer_bit_width() == m.metrics["circuit_max_bits"]
    assert np.abs(m.predict_clear(X) - y).max() < 0.1      # accuracy envelope (test_fixed_pipeline.py:77)
    with pytest.raises(ValueError, match="Unknown similarity type"):
        FHESimilarityModel(similarity_type="nope", verbose=False)._prepare_training_data(4)
    p = tmp_path / "m.pkl"
    m.save(str(p))
    data = pickle.load(open(p, "rb"))
    assert {"input_dim", "n_bits", "similarity_type", "metrics", "model_params"} <= set(data)
    m2 = FHESimilarityModel.load(str(p))
    assert not m2.compiled and np.array_equal(m2.predict_clear(X[:5]), m.predict_clear(X[:5]))


def test_seeded_generator_matches_reference_draw_order():
    """seed=S reproduces what the reference generates after np.random.seed(S) (fhe_similarity.py:41-58)."""
    from fhe_icp_b200 import FHESimilarityModel
    np.random.seed(5)
    n, d = 50, 16
    e1 = np.random.randn(n, d).astype(np.float32); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.random.randn(n, d).astype(np.float32); e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    mask = np.random.rand(n) > 0.5
    e2[mask] = e1[mask] + 0.2 * np.random.randn(mask.sum(), d)
    e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    X, y = FHESimilarityModel(input_dim=d, seed=5, verbose=False)._prepare_training_data(n)
    assert np.array_equal(X, e1 * e2) and np.array_equal(y, np.sum(e1 * e2, axis=1))


def test_estimator_members_callers_reach_through():
    from fhe_icp_b200 import LinearRegression, SGDRegressor
    rng = np.random.RandomState(0)
    X = rng.randn(200, 6).astype(np.float32)
    y = X @ np.array([1, -2, 0.5, 0, 3, 1], dtype=np.float32) + 0.25
    for est in (LinearRegression(n_bits=8), SGDRegressor(n_bits=8, max_iter=100, random_state=42)):
        est.fit(X, y)
        assert est.coef_.shape == (6,) and isinstance(est.intercept_, float)
        assert est.score(X, y) > 0.98
        est.compile(X[:20])
        assert est.fhe_circuit.graph.maximum_integer_bit_width() >= 8
        assert np.array_equal(est.predict(X[:3]), est.predict(X[:3], fhe="disable"))
        assert np.array_equal(est.predict(X[:3], fhe="simulate"), est.predict(X[:3]))
        with pytest.raises(ValueError):
            est.predict(X[:3], fhe="bogus")
        with pytest.raises(ValueError):
            est.predict(X[:3, :5])
    with pytest.raises(RuntimeError):
        LinearRegression().predict(X)


def test_weight_quantizer_both_regimes_and_two_output_decision():
    """float32 fits give a huge weight zero-point (two encrypted outputs, client-side
    correction), float64 fits degenerate to q_W = 1 (one output).  Both stay exact."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=42, verbose=False)
    X, y = m._prepare_training_data(1000)
    m.train(X.astype(np.float64), y.astype(np.float64))
    m.compile(X[:10])
    c = m.model.fhe_circuit
    assert not c.two_outputs and set(m.model.spec.q_weights.tolist()) == {1} and m.model.spec.weight_q.zero_point == 0
    assert c.lwe.n < 1100 and (c.lwe.n + 1) % 16 == 0
    # a hand-made spread of coefficients (asymmetric quantizer, zero-point far from 0)
    from fhe_icp_b200.quantization import QuantizedLinearSpec
    coef = 1.0 + 6e-6 * np.linspace(-0.5, 0.5, 128)
    spec = QuantizedLinearSpec.from_fit(coef, 0.0, X, 8)
    assert abs(spec.weight_q.zero_point) > 10 ** 6 and spec.q_weights.min() == -128 and spec.q_weights.max() == 127
    q = spec.input_q.quant(X[:50])
    ref = (spec.input_q.dequant(q) * spec.weight_q.dequant(spec.q_weights)).sum(axis=1)
    assert np.abs(spec.predict_clear(X[:50]) - ref).max() < 1e-3   # integer circuit == dequantized algebra


def test_parameter_selection_follows_noise_bound():
    from fhe_icp_b200.params import log2_sigma_for_dimension, select_lwe_params, z_score
    assert abs(log2_sigma_for_dimension(742) + 17.06) < 0.05 and abs(log2_sigma_for_dimension(2048) + 51.67) < 0.05
    assert abs(z_score(2 ** -40) - 7.15) < 0.05
    prev = 0
    for bits in (12, 16, 20, 24):
        p = select_lwe_params(bits, 2.0 ** 19)
        assert (p.n + 1) % 16 == 0 and p.stride % 2 == 0 and p.stride >= p.n + 1 and p.n >= prev
        prev = p.n
        # z * sigma * ||w|| < Delta / 2
        assert np.log2(z_score(p.p_error)) + p.log2_out_noise < p.shift - 1 - 64 + 1e-9
    with pytest.raises(ValueError, match="NoParametersFound"):
        select_lwe_params(62, 1.0)


def test_signed_bit_width():
    from fhe_icp_b200.quantization import signed_bit_width
    assert signed_bit_width(0, 255) == 8 and signed_bit_width(-128, 127) == 8 and signed_bit_width(-129, 0) == 9
    assert signed_bit_width(0, 0) == 1 and signed_bit_width(-1, 0) == 1 and signed_bit_width(-1, 1) == 2


def test_top_indices_equals_full_stable_sort():
    """The selection-based ranking equals the reference's 'filter, stable sort descending, [:k]' on ties, NaN-free
    ragged inputs, thresholds that empty the list and k larger than the collection."""
    from fhe_icp_b200.batch_operations import rank_results, top_indices
    rng = np.random.RandomState(0)
    for n in (0, 1, 5, 300, 5000):
        for trial in range(6):
            s = np.round(rng.randn(n), 1 if trial % 2 else 6)          # 1 decimal: many exact ties
            for k in (0, 1, 3, 10, n + 5):
                for thr in (-np.inf, -0.5, 0.5, 10.0):
                    keep = [i for i in range(n) if s[i] >= thr]
                    want = sorted(keep, key=lambda i: s[i], reverse=True)[:k]     # Python's stable sort, like the reference
                    assert top_indices(s, k, thr).tolist() == want
    ids = [f"d{i}" for i in range(7)]
    sc = np.array([0.5, 0.9, 0.5, 0.9, 0.1, 0.7, 0.5])
    assert rank_results(ids, sc, 4, 0.5) == [("d1", 0.9), ("d3", 0.9), ("d5", 0.7), ("d0", 0.5)]


def _reference_quantization_dataset(n_samples=500, dim=128):
    """The seeded generator of /root/reference/quantization_strategy.py:134-160 (seed 42, 256-d concatenation)."""
    np.random.seed(42)
    emb1 = np.random.randn(n_samples, dim).astype(np.float32)
    emb1 = emb1 / np.linalg.norm(emb1, axis=1, keepdims=True)
    emb2 = np.zeros_like(emb1)
    for i in range(n_samples):
        emb2[i] = emb1[i] + 0.1 * np.random.randn(dim) if i % 2 == 0 else np.random.randn(dim)
    emb2 = emb2 / np.linalg.norm(emb2, axis=1, keepdims=True)
    return np.hstack([emb1, emb2]), np.sum(emb1 * emb2, axis=1)


@pytest.mark.parametrize("n_bits,published", [(4, 12), (8, 20), (12, 28)])
def test_published_circuit_bit_widths(n_bits, published):
    """The one numeric fixture the reference publishes for this path: `Circuit Max Bits` 12 / 20 / 28 at 4 / 8 / 12-bit
    quantization (/root/reference/SESSION_REPORT.md:66-71), produced by quantization_strategy.py:34-59 on its seed-42
    dataset (SGDRegressor(max_iter=20, random_state=42), compile(X_train), graph.maximum_integer_bit_width())."""
    from fhe_icp_b200 import SGDRegressor
    X, y = _reference_quantization_dataset()
    X_train, y_train = X[:400], y[:400]
    model = SGDRegressor(n_bits=n_bits, max_iter=20, random_state=42)
    model.fit(X_train, y_train)
    circuit = model.compile(X_train)
    assert model.fhe_circuit.graph.maximum_integer_bit_width() == published
    # the sizing input stays separate: only encrypted values enter the parameter selection
    assert circuit.inputset_bits == published - 2 and circuit.guaranteed_bits >= circuit.inputset_bits


def test_quantized_circuit_obeys_affine_quantization_algebra():
    """Checks of the quantizer and of the integer circuit that do not go through oracle/oracle.py (whose restatement of
    SURVEY Appendix A is necessarily close to the product's): they follow from what affine quantization IS.
    (1) the calibration range maps onto the whole integer range, monotonically, and dequant(quant(x)) is within half a
    step of x inside the range; (2) the integer circuit equals the real-valued dot product of the DEQUANTIZED operands
    plus the intercept up to the rounding of the bias: |y - (X_deq @ W_deq + b)| <= s_X s_W / 2; (3) the exact rational
    value of the circuit (fractions, no floating point) agrees with (2)."""
    from fractions import Fraction
    from fhe_icp_b200.quantization import QuantizedLinearSpec, UniformQuantizer
    rng = np.random.default_rng(12)
    for n_bits in (2, 4, 8, 12):
        for trial in range(6):
            d = int(rng.integers(1, 40))
            X = rng.normal(scale=rng.choice([1e-3, 0.1, 3.0]), size=(50, d)) + rng.normal()
            coef = rng.normal(scale=rng.choice([1e-2, 1.0, 7.0]), size=d)
            b = float(rng.normal())
            q = UniformQuantizer.from_values(X, n_bits)
            lo, hi = X.min(), X.max()
            assert q.quant(np.array([lo]))[0] == q.qmin and q.quant(np.array([hi]))[0] == q.qmax
            xs = np.sort(rng.uniform(lo, hi, size=400))
            qx = q.quant(xs)
            assert np.all(np.diff(qx) >= 0)
            assert np.max(np.abs(q.dequant(qx) - xs)) <= q.scale * (0.5 + 1e-9)
            spec = QuantizedLinearSpec.from_fit(coef, b, X, n_bits)
            Xt = rng.uniform(lo, hi, size=(20, d))
            qX = spec.input_q.quant(Xt)
            X_deq = spec.input_q.dequant(qX)
            W_deq = spec.weight_q.dequant(spec.q_weights)
            y = spec.predict_clear(Xt)
            s = spec.input_q.scale * spec.weight_q.scale
            assert np.max(np.abs(y - (X_deq @ W_deq + b))) <= s * 0.5 + 1e-9 * max(1.0, np.abs(y).max())
            # exact rational evaluation of the first row
            sx, sw = Fraction(spec.input_q.scale), Fraction(spec.weight_q.scale)
            exact = sum(sx * (int(a) - spec.input_q.zero_point) * sw * (int(w) - spec.weight_q.zero_point)
                        for a, w in zip(qX[0], spec.q_weights))
            assert abs(float(exact) + b - y[0]) <= float(sx * sw) * 0.5 + 1e-9 * max(1.0, abs(y[0]))


def test_cli_clear_mode_on_cpu_and_loud_failure_of_execute_without_a_gpu(tmp_path, capsys):
    """The three sub-commands the reference's CLI keeps (`/root/reference/fhe_cli.py:106-210,327-346`) in `--fhe disable` --
    the clear quantized model, which is what the reference's own CLI evaluates (`batch_operations.py:233,276`) -- need no
    GPU: same arguments, same printed lines, ids file, metadata, threshold and top-k.  `--fhe execute` without a CUDA device
    must fail loudly (exit status 1, as the reference's handler does on any error) and never fall back to the CPU."""
    import json
    import torch
    from fhe_icp_b200.fhe_cli import main
    docs = [{"text": "quantum computing with qubits", "id": "q1", "metadata": {"tag": "physics"}},
            {"text": "quantum error correction", "id": "q2"}, {"text": "cooking pasta recipes", "id": "c1"}, "cooking garlic"]
    (tmp_path / "docs.json").write_text(json.dumps(docs))
    sd = str(tmp_path / "store")
    assert main(["--storage-dir", sd, "--fhe", "disable", "encrypt-batch", str(tmp_path / "docs.json"),
                 "-o", str(tmp_path / "ids.json")]) == 0
    ids = json.loads((tmp_path / "ids.json").read_text())
    assert ids[:3] == ["q1", "q2", "c1"] and len(ids) == 4
    assert "Encrypted 4 documents successfully!" in capsys.readouterr().out
    assert main(["--storage-dir", sd, "--fhe", "disable", "compare", "q1", "q2"]) == 0
    out = capsys.readouterr().out
    assert "Document 1: q1" in out and "Similarity score:" in out and "Interpretation: Very similar" in out
    assert main(["--storage-dir", sd, "--fhe", "disable", "compare", "q1", "c1"]) == 0
    far = capsys.readouterr().out
    score = lambda s: float([l for l in s.splitlines() if "Similarity score" in l][0].split(":")[1])   # noqa: E731
    assert score(far) < score(out)
    assert main(["--storage-dir", sd, "--fhe", "disable", "search", "quantum supremacy", "--top-k", "3"]) == 0
    s = capsys.readouterr().out
    assert "Found 2 similar documents" in s and "1. q" in s and "Metadata: {'tag': 'physics'}" in s
    assert main(["--storage-dir", sd, "--fhe", "disable", "search", "quantum supremacy", "--top-k", "1",
                 "--min-similarity", "0.0"]) == 0
    assert "Found 1 similar documents" in capsys.readouterr().out
    assert main(["--storage-dir", sd, "--fhe", "disable", "search", "quantum supremacy", "--min-similarity", "0.9999"]) == 0
    assert "No similar documents found." in capsys.readouterr().out
    if not torch.cuda.is_available():
        # compare: the reference's handler prints the error itself and returns (`fhe_cli.py:180-181`); search has no handler of
        # its own, so the error reaches main(), which reports it and exits 1 (`fhe_cli.py:384-390`)
        assert main(["--storage-dir", sd, "compare", "q1", "q2"]) == 0
        o = capsys.readouterr().out
        assert "no CPU fallback" in o and "Similarity score" not in o
        assert main(["--storage-dir", sd, "search", "quantum supremacy"]) == 1
        c = capsys.readouterr()
        assert "no CPU fallback" in c.err and "Found" not in c.out


def test_seed_policy_secrets_are_fresh_public_parts_carry_no_secret_and_ids_never_repeat():
    """The seed policy of randomness.py (what Concrete's CSPRNG key generation gives the reference for free): without explicit
    seeds every model draws its own key / noise / mask seeds and its own ciphertext-id origin; what travels to an evaluator
    (a SeededCiphertexts object, the evaluator-side C signatures) names the public mask seed only; the id allocator is
    monotonic and hands out disjoint ranges; explicit seeds reproduce."""
    import inspect
    from pathlib import Path
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200.fhe_similarity import SeededCiphertexts
    from fhe_icp_b200.randomness import CiphertextIds, seed_or_fresh

    def circuit(**kw):
        m = FHESimilarityModel(input_dim=16, n_bits=8, seed=1, verbose=False, **kw)      # `seed` seeds the data only
        X, _ = m.train(n_samples=60)
        m.compile(X[:10])
        return m.model.fhe_circuit

    a, b = circuit(), circuit()
    secrets_a = {a.key_seed, a.noise_seed, a.enc_seed}
    assert len(secrets_a) == 3 and not secrets_a & {b.key_seed, b.noise_seed, b.enc_seed}
    assert a.ct_counter != b.ct_counter and a.ct_counter < 2 ** 62
    assert np.array_equal(a.spec.q_weights, b.spec.q_weights)          # same data seed: same public model
    c1, c2 = circuit(key_seed=5, noise_seed=6, enc_seed=7, ct_start=0), circuit(key_seed=5, noise_seed=6, enc_seed=7, ct_start=0)
    assert (c1.key_seed, c1.noise_seed, c1.enc_seed, c1.ct_counter) == (5, 6, 7, 0) == (c2.key_seed, c2.noise_seed, c2.enc_seed, c2.ct_counter)
    # the compressed ciphertext object and the evaluator-side entry points carry the public mask seed only
    assert set(inspect.signature(SeededCiphertexts.__init__).parameters) == {"self", "bodies", "enc_seed", "ct_base"}
    header = (Path(__file__).resolve().parent.parent / "include" / "fhe_b200.h").read_text()
    for fn in ("fhe_b200_similarity_run_seeded", "fhe_b200_similarity_run_seeded_push", "fhe_b200_similarity_create_evaluator"):
        decl = header[header.index(fn + "("):]
        decl = decl[: decl.index(";")]
        assert "noise_seed" not in decl and "key_seed" not in decl, fn
    # ids: monotonic, disjoint, random origin by default
    ids = CiphertextIds(10)
    assert [ids.take(3), ids.take(0), ids.take(5), ids.take(1)] == [10, 13, 13, 18] and ids.next == 19
    assert CiphertextIds().next != CiphertextIds().next
    assert c1.next_ct_base(4) == 0 and c1.next_ct_base(2) == 4 and c1.ct_counter == 6
    assert seed_or_fresh(None) != seed_or_fresh(None) and seed_or_fresh(-1) == 2 ** 64 - 1
