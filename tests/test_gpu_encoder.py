"""GPU: the BERT-family encoders of csrc/encoder.cuh against the fp32 PyTorch models.

The reference's checkpoints (all-MiniLM-L6-v2, ms-marco-TinyBERT-L2-v2) cannot be downloaded here, so the oracle is
``transformers``' own BertModel / BertForSequenceClassification with the SAME architecture and seeded random
weights (perturbed away from the initialiser so LayerNorm gains, biases and the position / type tables all
matter), evaluated in fp32 on the CPU exactly the way sentence-transformers drives it: padded batch + attention
mask -> last_hidden_state -> mean pooling over the mask -> L2 normalise (semantic_search.py:211-222), and
[CLS] pooler -> 1-logit classifier for the cross-encoder (hybrid_search.py:296-297).

Tolerance (floating point, stated here): normwise relative error of the pooled vector <= 1e-5 against the fp32
PyTorch result — and, as a sanity check of that bar, <= 5e-6 against an fp64 evaluation of the same model
(PyTorch-fp32's own error against fp64 is ~1e-6 on these inputs).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "LayerNorm.weight" in name:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "embeddings" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) / (p.shape[-1] ** 0.5))
    return model.eval()


def _batch(seed, n, lo, hi, vocab, pair=False):
    rng = np.random.default_rng(seed)
    ids, types = [], []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        x = rng.integers(1000, vocab, L).tolist()
        x[0], x[-1] = 101, 102                                   # [CLS] ... [SEP]
        t = [0] * L
        if pair and L > 4:
            cut = int(rng.integers(2, L - 1))
            x[cut] = 102
            t = [0] * (cut + 1) + [1] * (L - cut - 1)
        ids.append(x); types.append(t)
    return ids, types


def _pad(lists, fill=0):
    S = max(len(x) for x in lists)
    a = np.full((len(lists), S), fill, np.int64)
    m = np.zeros((len(lists), S), np.int64)
    for i, x in enumerate(lists):
        a[i, :len(x)] = x
        m[i, :len(x)] = 1
    return torch.from_numpy(a), torch.from_numpy(m)


def _st_pipeline(bert, ids, dtype):
    """sentence-transformers: Transformer -> Pooling(mean) -> Normalize."""
    x, mask = _pad(ids)
    with torch.no_grad():
        h = bert(input_ids=x, attention_mask=mask).last_hidden_state.to(dtype)
    m = mask.unsqueeze(-1).to(dtype)
    pooled = (h * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
    return torch.nn.functional.normalize(pooled, p=2, dim=1).numpy()


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b, axis=-1).max() / max(1e-30, np.linalg.norm(b, axis=-1).min()))


@pytest.fixture(scope="module")
def minilm():
    from rag_search_engine_b200.encoder import MINILM_L6_CONFIG as C
    cfg = transformers.BertConfig(vocab_size=C["vocab_size"], hidden_size=C["hidden"], num_hidden_layers=C["layers"],
                                  num_attention_heads=C["heads"], intermediate_size=C["intermediate"],
                                  max_position_embeddings=C["max_positions"], type_vocab_size=C["type_vocab"],
                                  layer_norm_eps=C["ln_eps"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    return _randomise(transformers.BertModel(cfg, add_pooling_layer=False), 1)


def test_minilm_query_encoder_matches_fp32_pytorch(minilm, fresh_index):
    from rag_search_engine_b200.encoder import GpuSentenceEncoder, config_from_hf
    enc = GpuSentenceEncoder(fresh_index, minilm.state_dict(), config_from_hf(minilm.config))
    assert enc.get_sentence_embedding_dimension() == 384
    for seed, n, lo, hi in ((3, 67, 3, 24), (4, 5, 100, 200), (5, 1, 2, 2), (6, 9, 250, 256)):
        ids, _ = _batch(seed, n, lo, hi, 30522)
        got = enc.encode_ids(ids)
        ref32 = _st_pipeline(minilm, ids, torch.float32)
        assert got.shape == ref32.shape == (n, 384)
        e = _relerr(got, ref32)
        assert e <= 1e-5, f"pooled vectors differ from the fp32 PyTorch model by {e:.2e} (batch {seed})"
        assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - 1.0).max() < 1e-6
    # the bar in perspective: our error against an fp64 evaluation vs PyTorch-fp32's own — for BOTH GEMM paths
    # (0 = tcgen05 kind::tf32 with the 3-pass hi/lo split, the default; 1 = fp32 SIMT)
    ids, _ = _batch(7, 24, 3, 40, 30522)
    ref64 = _st_pipeline(minilm.double(), ids, torch.float64)
    minilm.float()
    theirs = _relerr(_st_pipeline(minilm, ids, torch.float32), ref64)
    got = {}
    for mode, name in ((0, "tcgen05 3xTF32"), (1, "fp32 SIMT")):
        fresh_index.encoder_set_mode(enc.slot, mode)
        got[mode] = enc.encode_ids(ids)
        ours = _relerr(got[mode], ref64)
        print(f"error vs fp64: librse {name} {ours:.2e}, PyTorch fp32 {theirs:.2e}")
        assert ours <= 5e-6
        assert _relerr(got[mode], _st_pipeline(minilm, ids, torch.float32)) <= 1e-5
    assert _relerr(got[0], got[1]) <= 5e-6
    fresh_index.encoder_set_mode(enc.slot, 0)
    # sequences longer than max_seq_length are truncated like the tokenizer would; an empty one is refused
    long_ids, _ = _batch(8, 2, 300, 300, 30522)
    assert np.array_equal(enc.encode_ids(long_ids), enc.encode_ids([x[:256] for x in long_ids]))
    from rag_search_engine_b200._lib import RseError
    with pytest.raises(RseError, match="empty"):
        enc.encode_ids([[101, 102], []])


def test_tinybert_cross_encoder_matches_fp32_pytorch(fresh_index):
    from rag_search_engine_b200.encoder import TINYBERT_L2_CONFIG as C, GpuCrossEncoder, config_from_hf
    cfg = transformers.BertConfig(vocab_size=C["vocab_size"], hidden_size=C["hidden"], num_hidden_layers=C["layers"],
                                  num_attention_heads=C["heads"], intermediate_size=C["intermediate"],
                                  max_position_embeddings=C["max_positions"], type_vocab_size=C["type_vocab"],
                                  layer_norm_eps=C["ln_eps"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                  num_labels=1)
    model = _randomise(transformers.BertForSequenceClassification(cfg), 2)
    ce = GpuCrossEncoder(fresh_index, model.state_dict(), config_from_hf(model.config))
    for seed, n, lo, hi in ((11, 20, 12, 90), (12, 3, 400, 512), (13, 1, 5, 5)):
        ids, types = _batch(seed, n, lo, hi, 30522, pair=True)
        x, mask = _pad(ids)
        tt, _ = _pad(types)
        with torch.no_grad():
            ref = model(input_ids=x, attention_mask=mask, token_type_ids=tt).logits[:, 0].numpy()
        got = ce.predict_ids(ids, types)
        assert got.shape == (n,)
        err = np.abs(got.astype(np.float64) - ref) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= 1e-5, f"cross-encoder logits off by {err.max():.2e}"
        assert np.argsort(-got, kind="stable").tolist() == np.argsort(-ref, kind="stable").tolist()
    ce.activation = "sigmoid"
    sg = ce.predict_ids(ids, types)
    assert np.allclose(sg, 1.0 / (1.0 + np.exp(-ref)), atol=1e-6)


def test_text_in_hybrid_keeps_the_query_vectors_on_the_device(minilm):
    """rse_encode_dev -> rse_hybrid_stage_dev -> run == rse_encode (host) -> rse_hybrid, bit for bit; and the
    cross-encoder plug of HybridSearch.rrf_search reorders by its scores (hybrid_search.py:279-312)."""
    from rag_search_engine_b200 import HybridSearch, _lib, synth
    from rag_search_engine_b200.encoder import GpuSentenceEncoder, config_from_hf
    se = synth.synth_embeddings(3000, seed=21, device="cpu")
    bm = synth.synth_bm25(3000, 2000, seed=21, mean_len=30, sd_len=10)
    idx = _lib.Index(0)
    try:
        idx.load_embeddings(se.emb.numpy(), movie_idx=se.movie_of_chunk.numpy())
        idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
        enc = GpuSentenceEncoder(idx, minilm.state_dict(), config_from_hf(minilm.config))
        T = len(bm.df)
        names = [f"t{i}" for i in range(T)]
        hs = HybridSearch.from_loaded(idx, dict(zip(names, range(T))), se.movie_ids, se.movie_ids,
                                      registry_key="r02-textin-test")
        nq = 40
        tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=22)
        token_lists = [[names[t] if t >= 0 else "<oov>" for t in terms[tok_indptr[i]:tok_indptr[i + 1]]] for i in range(nq)]
        ids, _ = _batch(23, nq, 3, 20, 30522)
        got = hs.rrf_search_texts(token_lists, ids, enc, k=60, limit=10, as_arrays=True)
        vecs = enc.encode_ids(ids)
        want = hs.rrf_search_batch(token_lists, vecs, k=60, limit=10, as_arrays=True)
        assert all((a.view(np.uint8) == b.view(np.uint8)).all() for a, b in zip(got, want))
        assert int(want[4].sum()) > 0
    finally:
        idx.close()


def test_bulk_encode_is_sliced_and_equal(minilm, fresh_index):
    """Index-build sized input (semantic_search.py:199-206): slices of the token budget give the same vectors."""
    from rag_search_engine_b200.encoder import GpuSentenceEncoder, config_from_hf
    enc = GpuSentenceEncoder(fresh_index, minilm.state_dict(), config_from_hf(minilm.config))
    ids, _ = _batch(31, 300, 3, 30, 30522)
    whole = enc.encode_ids(ids)
    sliced = enc.encode_ids(ids, max_tokens_per_call=500)
    assert np.array_equal(whole, sliced) and whole.shape == (300, 384)


def _word_ids(text, lo=1000, span=29000):
    import zlib
    return [101] + [lo + zlib.crc32(w.encode()) % span for w in text.lower().split()] + [102]


def test_mirrors_take_the_gpu_encoders_through_the_reference_seams(minilm, tmp_path):
    """The reference calls ``self.model.encode(texts)`` (semantic_search.py:221) and
    ``CrossEncoder(...).predict(pairs)`` then sorts by (cross_encoder_score, score) (hybrid_search.py:296-309):
    SemanticSearch(encoder=GpuSentenceEncoder) and HybridSearch(cross_encoder=GpuCrossEncoder) go through exactly
    those seams, with the models on the GPU."""
    from rag_search_engine_b200 import HybridSearch, runtime, store
    from rag_search_engine_b200.encoder import TINYBERT_L2_CONFIG as C, GpuCrossEncoder, GpuSentenceEncoder, config_from_hf
    from rag_search_engine_b200.textutil import whitespace_tokenizer
    rng = np.random.default_rng(5)
    words = [f"w{i}" for i in range(60)]
    docs = [{"id": 7 + 3 * i, "title": " ".join(rng.choice(words, 2)).title(),
             "description": " ".join(rng.choice(words, 12)) + ". " + " ".join(rng.choice(words, 9)) + "."} for i in range(120)]
    db = tmp_path / "enc.db"
    idx = runtime.acquire(db, 0)                       # the handle the mirrors will share
    try:
        enc = GpuSentenceEncoder(idx, minilm.state_dict(), config_from_hf(minilm.config),
                                 tokenizer=lambda texts: [_word_ids(t) for t in texts])
        store.write_reference_db(db, docs, whitespace_tokenizer, embed=enc.encode)        # chunk embeddings by the GPU encoder
        cfg = transformers.BertConfig(vocab_size=C["vocab_size"], hidden_size=C["hidden"], num_hidden_layers=C["layers"],
                                      num_attention_heads=C["heads"], intermediate_size=C["intermediate"],
                                      max_position_embeddings=C["max_positions"], type_vocab_size=C["type_vocab"],
                                      layer_norm_eps=C["ln_eps"], num_labels=1)
        ce_model = _randomise(transformers.BertForSequenceClassification(cfg), 9)

        def pair_tok(pairs):
            ids, tts = [], []
            for q, d in pairs:
                a, b = _word_ids(q), _word_ids(d)[1:]
                ids.append(a + b); tts.append([0] * len(a) + [1] * len(b))
            return ids, tts
        ce = GpuCrossEncoder(idx, ce_model.state_dict(), config_from_hf(ce_model.config), pair_tokenizer=pair_tok)
        hs = HybridSearch(docs_path=None, db_path=db, tokenizer=whitespace_tokenizer, encoder=enc, cross_encoder=ce)
        try:
            q = "w3 w17 w5"
            sem = hs.semantic.query_top_k(q, k=5)
            assert len(sem) == 5 and sem == sorted(sem, key=lambda h: h["distance"])
            # the query vector the mirror used == the encoder's own output (and a unit vector)
            v = hs.semantic.generate_embedding(q)
            assert v.shape == (1, 384) and abs(float(np.linalg.norm(v[0])) - 1.0) < 1e-6
            seen = []
            orig_predict = ce.predict
            ce.predict = lambda pairs, **kw: (seen.append([list(p) for p in pairs]), orig_predict(pairs, **kw))[1]
            rr = hs.rrf_search(q, k=60, limit=8, rerank_method="cross_encoder")
            ce.predict = orig_predict
            # the reranker saw the WHOLE fused union (the reference truncates after reranking, :312), in RRF order
            pairs = seen[0]
            assert len(rr) == 8 and 8 <= len(pairs) <= 16 and all(p[0] == q for p in pairs)
            text_to_id = {f"{d['title']} - {d['description']}": d["id"] for d in docs}
            pair_ids = [text_to_id[p[1]] for p in pairs]
            # the fp32 PyTorch cross-encoder is the oracle for the scores; the reference's rule orders them (:304-309)
            ids, tts = pair_tok(pairs)
            x, mask = _pad(ids)
            tt, _ = _pad(tts)
            with torch.no_grad():
                ref = ce_model(input_ids=x, attention_mask=mask, token_type_ids=tt).logits[:, 0].numpy()
            by_id = dict(zip(pair_ids, ref))
            for d in rr:
                assert abs(d["cross_encoder_score"] - by_id[d["id"]]) <= 1e-5 * max(1.0, abs(by_id[d["id"]]))
            want = [pair_ids[i] for i in np.argsort(-ref, kind="stable")[:8]]
            assert [d["id"] for d in rr] == want
            assert all(d["rrf_rank"] == pair_ids.index(d["id"]) + 1 for d in rr)
        finally:
            hs.close()
    finally:
        runtime.release(db, 0)


def test_encoder_abi_error_paths(fresh_index):
    """Status codes instead of undefined behaviour: wrong tensor sizes / names, missing tensors, use before
    finalize, unsupported shapes, over-long sequences."""
    from rag_search_engine_b200._lib import RseError
    from rag_search_engine_b200.encoder import random_state_dict
    idx = fresh_index
    cfg = dict(vocab_size=500, hidden=128, layers=1, heads=2, intermediate=512, max_positions=16, type_vocab=2, ln_eps=1e-12)
    with pytest.raises(RseError, match="create"):
        idx._check(idx._L.rse_encoder_finalize(idx._h, 0))
    with pytest.raises(RseError, match="unsupported"):
        idx.encoder_create(0, head=0, **{**cfg, "hidden": 100})
    with pytest.raises(RseError, match="bad arguments"):
        idx.encoder_create(5, head=0, **cfg)
    idx.encoder_create(0, head=0, **cfg)
    sd = random_state_dict(cfg, seed=1)
    with pytest.raises(RseError, match="element count"):
        idx.encoder_set_tensor(0, "embeddings.LayerNorm.weight", np.zeros(64, np.float32))
    with pytest.raises(RseError, match="unknown tensor|not a BERT"):
        idx.encoder_set_tensor(0, "encoder.layer.0.attention.self.nope.weight", np.zeros(4, np.float32))
    with pytest.raises(RseError, match="layer index"):
        idx.encoder_set_tensor(0, "encoder.layer.3.output.dense.bias", np.zeros(128, np.float32))
    names = list(sd)
    for n in names[:-1]:
        idx.encoder_set_tensor(0, "0.auto_model." + n, sd[n])             # a wrapper prefix is ignored
    with pytest.raises(RseError, match="tensors were set"):
        idx.encoder_finalize(0)
    with pytest.raises(RseError, match="not finalized"):
        idx.encode(0, np.array([101, 102], np.int32), np.array([0, 2], np.int32))
    idx.encoder_set_tensor(0, names[-1], sd[names[-1]])
    idx.encoder_finalize(0)
    out = idx.encode(0, np.array([101, 7, 102, 101, 102], np.int32), np.array([0, 3, 5], np.int32))
    assert out.shape == (2, 128) and np.isfinite(out).all()
    with pytest.raises(RseError, match="max_positions"):
        idx.encode(0, np.arange(20, dtype=np.int32), np.array([0, 20], np.int32))
    with pytest.raises(ValueError):
        idx.encode(0, np.arange(5, dtype=np.int32), np.array([0, 4], np.int32))          # cu_seqlens does not cover the ids


def test_fallback_build_is_explicit_and_marked(minilm, tmp_path):
    """Index build stays on the reference (north_star).  Without the reference package the mirror refuses to build
    unless asked (fallback_build=True), and the database it then writes says that sqlite-vec cannot open it."""
    import json
    import sqlite3
    from rag_search_engine_b200 import SemanticSearch, runtime
    from rag_search_engine_b200.encoder import GpuSentenceEncoder, config_from_hf
    try:
        import rag_search_engine.utils.semantic_search  # noqa: F401
        pytest.skip("the reference package is importable here: its own build is used")
    except ImportError:
        pass
    docs = {"movies": [{"id": 11 + i, "title": f"T{i} w{i % 5}", "description": f"w{i} w{i + 1}. w{i % 7} again."} for i in range(40)]}
    p = tmp_path / "movies.json"
    p.write_text(json.dumps(docs))
    db = tmp_path / "fb.db"
    idx = runtime.acquire(db, 0)
    try:
        enc = GpuSentenceEncoder(idx, minilm.state_dict(), config_from_hf(minilm.config),
                                 tokenizer=lambda texts: [_word_ids(t) for t in texts])
        with pytest.raises(RuntimeError, match="fallback_build=True"):
            SemanticSearch(docs_path=p, db_path=db, encoder=enc)
        ss = SemanticSearch(docs_path=p, db_path=db, encoder=enc, fallback_build=True)     # bulk embed on the GPU encoder
        try:
            hits = ss.query_top_k("w3 w4", k=3)
            assert len(hits) == 3 and all(h["chunk"] for h in hits)
            (note,) = sqlite3.connect(db).execute("SELECT value FROM rse_meta WHERE key='vec0_shadow_writer'").fetchone()
            assert "not openable by sqlite-vec" in note
        finally:
            ss.close()
    finally:
        runtime.release(db, 0)
