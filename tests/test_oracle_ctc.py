"""Pin the CTC oracle (oracle/ctc_oracle.c) with independent cross-checks.

The reference ships no golden vectors for this path ("parity unpinned"); what we can check:
  * loss/grad  vs torch.nn.functional.ctc_loss + autograd in fp64 on log_softmax(log(p+eps))
  * Sum_k grad_k == 0 per frame, infeasible-label error
  * beam search (unbounded width, merge_repeated=False) == brute-force most probable labelling
  * merge_repeated artefact of the reference ("cellist" -> "celist")
  * greedy == collapse of the argmax path
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import ctc_oracle as O


def _rand_probs(rng, B, T, V, scale=3.0):
    z = rng.standard_normal((B, T, V)).astype(np.float32) * scale
    z -= z.max(-1, keepdims=True)
    p = np.exp(z)
    return (p / p.sum(-1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("T,V,maxL", [(50, 38, 23), (64, 38, 23), (25, 96, 12), (7, 5, 3)])
def test_loss_grad_vs_torch(T, V, maxL):
    rng = np.random.default_rng(T * 1000 + V)
    B = 12
    probs = _rand_probs(rng, B, T, V)
    lens = rng.integers(1, min(maxL, T // 2) + 1, B).astype(np.int32)
    lens[0] = 1
    labels = np.full((B, maxL), V - 1, np.int32)
    for b in range(B):
        lab = rng.integers(0, V - 1, lens[b])
        if b % 3 == 0 and lens[b] >= 2:
            lab[1] = lab[0]  # repeated chars
        labels[b, :lens[b]] = lab
    in_len = np.full(B, T, np.int32)
    in_len[1] = T - 3  # shorter sequence: trailing frames get zero grad
    loss, grad = O.ctc_loss_grad(probs, labels, lens, in_len)

    u = torch.log(torch.tensor(probs, dtype=torch.float64) + 1e-7).requires_grad_(True)
    lsm = torch.log_softmax(u, -1).transpose(0, 1)
    tl = torch.nn.functional.ctc_loss(lsm, torch.tensor(labels, dtype=torch.long), torch.tensor(in_len, dtype=torch.long),
                                      torch.tensor(lens, dtype=torch.long), blank=V - 1, reduction="none")
    tl.sum().backward()
    np.testing.assert_allclose(loss, tl.detach().numpy(), rtol=2e-5, atol=1e-4)
    np.testing.assert_allclose(grad, u.grad.numpy(), rtol=0, atol=2e-4)  # fp32 log-space at |log p|~100: ulp 7.6e-6
    # per-frame gradient sums to zero inside the sequence, and is exactly 0 after it
    assert np.abs(grad.sum(-1)).max() < 5e-4
    assert np.all(grad[1, T - 3:] == 0)


def test_softmax_of_log_identity():
    # TF re-softmaxes u=log(p+eps): q = (p+eps)/(1+V*eps)  
    rng = np.random.default_rng(0)
    p = _rand_probs(rng, 1, 4, 38)[0].astype(np.float64)
    p /= p.sum(-1, keepdims=True)
    u = np.log(p + 1e-7)
    q = np.exp(u - u.max(-1, keepdims=True))
    q /= q.sum(-1, keepdims=True)
    np.testing.assert_allclose(q, (p + 1e-7) / (1 + 38e-7), rtol=1e-12)


def test_infeasible_labels_raise():
    probs = _rand_probs(np.random.default_rng(1), 1, 3, 5)
    labels = np.array([[1, 1, 2]], np.int32)  # needs 4 frames (repeat) > 3
    with pytest.raises(ValueError, match="Not enough time"):
        O.ctc_loss_grad(probs, labels, np.array([3]), np.array([3]))


def _collapse(path, blank):
    out, prev = [], -1
    for k in path:
        if k != blank and k != prev:
            out.append(int(k))
        prev = k
    return tuple(out)


def test_beam_unbounded_equals_bruteforce():
    rng = np.random.default_rng(7)
    T, V = 5, 4
    blank = V - 1
    for trial in range(30):
        probs = _rand_probs(rng, 1, T, V, scale=1.5)
        # brute force over all V^T paths on the distribution the decoder sees (p+eps; un-normalised is fine)
        score = {}
        pe = probs[0].astype(np.float64) + 1e-7
        for path in itertools.product(range(V), repeat=T):
            pr = np.prod([pe[t, k] for t, k in enumerate(path)])
            key = _collapse(path, blank)
            score[key] = score.get(key, 0.0) + pr
        best = max(score, key=score.get)
        out, n, lp = O.beam(probs, beam_width=4096, merge_repeated=False)
        assert tuple(out[0, :n[0]]) == best, (trial, best, out[0])


def test_beam_merge_repeated_artefact():
    # confident "a, blank, a" -> merge_repeated=True collapses the emitted double letter (reference README.md:45)
    V = 4
    blank = V - 1
    T = 3
    p = np.full((1, T, V), 0.01, np.float32)
    for t, k in enumerate([0, blank, 0]):
        p[0, t, k] = 0.97
    out, n, _ = O.beam(p, beam_width=10, merge_repeated=True)
    assert list(out[0, :n[0]]) == [0]
    out, n, _ = O.beam(p, beam_width=10, merge_repeated=False)
    assert list(out[0, :n[0]]) == [0, 0]


def test_beam_onehot_equals_collapse_and_greedy():
    rng = np.random.default_rng(3)
    B, T, V = 16, 25, 38
    paths = rng.integers(0, V, (B, T))
    p = np.full((B, T, V), 1e-4, np.float32)
    for b in range(B):
        p[b, np.arange(T), paths[b]] = 1.0 - 1e-4 * (V - 1)
    g, gn, gs = O.greedy(p)
    bo, bn, _ = O.beam(p, beam_width=10, merge_repeated=False)
    for b in range(B):
        ref = _collapse(paths[b], V - 1)
        assert tuple(g[b, :gn[b]]) == ref
        assert tuple(bo[b, :bn[b]]) == ref
        assert np.all(g[b, gn[b]:] == -1)


def test_beam_parent_sequential_formulation():
    """The formulation the CUDA kernel uses (parents visited best-first, closed-form 'blocked survivor' test, top-W
    merge per parent; oracle/beam_reference_py.py) equals the literal TF sequential insert/evict restatement --
    including on near-uniform inputs, where TF's "Deactivate child" side effect changes the result and a naive
    global top-W would NOT match (found on the first GPU run)."""
    from oracle.beam_reference_py import beam_parent_sequential
    for (B, T, V, sc, W, seed) in [(12, 25, 96, 3.0, 10, 11), (8, 52, 38, 6.0, 10, 12), (16, 66, 38, 1.0, 10, 178),
                                   (12, 40, 38, 0.5, 10, 2), (12, 30, 20, 1.5, 5, 3), (8, 30, 12, 1.0, 32, 6), (8, 20, 8, 0.7, 4, 7)]:
        probs = _rand_probs(np.random.default_rng(seed), B, T, V, sc)
        out, n, _ = O.beam(probs, beam_width=W, merge_repeated=True)
        for b in range(B):
            assert beam_parent_sequential(probs[b], W, True) == list(out[b, :n[b]]), (B, T, V, sc, W, seed, b)


def test_beam_kernel_shortcuts_are_exact():
    """The short cuts of ctc_beam_kernel's fast path (csrc/ctc.cu, beam width <= 16) -- the two pruning floors of the per-step candidate list,
    the blocking count taken on that list with its completeness test, the number of insertions per parent in closed form, the survivors'
    (total desc, slot desc) order -- restated in Python (oracle/beam_reference_py.py::beam_parent_sequential_fast) and checked against the
    literal TF restatement in ctc_oracle.c on far more shapes than the GPU tests run: flat to saturated inputs, vocabularies around the
    32-lane boundary, widths 1..16.  Every short cut must actually have been taken."""
    from oracle import beam_reference_py as R
    R.STATS.clear()
    rng0 = np.random.default_rng(2024)
    for i in range(48):
        V = int(rng0.choice([5, 12, 33, 34, 38, 65, 96, 97, 130]))
        W = int(rng0.choice([1, 2, 3, 5, 8, 10, 12, 16]))
        T = int(rng0.integers(1, 32))
        sc = float(rng0.choice([0.2, 0.5, 1.0, 2.0, 3.0, 6.0, 12.0]))
        probs = _rand_probs(np.random.default_rng(500 + i), 12, T, V, sc)
        out, n, _ = O.beam(probs, beam_width=W, merge_repeated=True)
        for b in range(probs.shape[0]):
            assert R.beam_parent_sequential_fast(probs[b], W, True) == list(out[b, :n[b]]), (T, V, sc, W, i, b)
    st = R.STATS
    assert st["floor1"] > 1000 and st["floor2"] > 500 and st["chunked"] > 50 and st["block_tests"] > 1000 and st["block_fallback"] > 5 and st["inserted"] > 10000, dict(st)


def test_threaded_driver_matches():
    probs = _rand_probs(np.random.default_rng(5), 37, 25, 96)
    a = O.beam(probs)
    b = O.beam_threaded(probs, 4)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def _tf_upstream():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tf_upstream_ctc.json")))


def test_oracle_reproduces_tf_upstream_ctc_loss_vectors():
    """TensorFlow's own ctc_loss_op_test.py::testBasic known answers (recalled constants, tests/golden/make_tf_upstream_golden.py): the C
    restatement of TF CTCLoss gives both losses and all 60 gradient entries to the 6 printed digits -- this pins the TRAINING half of the
    oracle (loss + gradient wrt the logits) on upstream-published vectors."""
    g = _tf_upstream()["ctc_loss"]
    probs = np.asarray(g["input_prob_matrix"], np.float32)
    maxL = max(len(t) for t in g["targets"])
    labels = np.full((2, maxL), g["depth"] - 1, np.int32)
    for b, t in enumerate(g["targets"]):
        labels[b, :len(t)] = t
    lens = [len(t) for t in g["targets"]]
    loss, grad = O.ctc_loss_grad(probs, labels, lens, [g["seq_len"]] * 2, eps=0.0)     # TF feeds log(p) as logits: softmax(log p) = p
    np.testing.assert_allclose(loss, g["loss"], rtol=0, atol=1.5e-5)
    np.testing.assert_allclose(grad, np.asarray(g["gradient"], np.float32), rtol=0, atol=1.5e-6)
    # the Keras path adds epsilon()=1e-7 before the log (utils.py:103 -> K.ctc_batch_cost): a 1e-6-level perturbation of the same numbers
    loss_k, grad_k = O.ctc_loss_grad(probs, labels, lens, [g["seq_len"]] * 2, eps=1e-7)
    np.testing.assert_allclose(loss_k, g["loss"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(grad_k, np.asarray(g["gradient"], np.float32), rtol=0, atol=5e-6)


def test_oracle_reproduces_tf_upstream_decoder_vectors():
    """TF ctc_decoder_ops_test.py::testCTCGreedyDecoder / testCTCDecoderBeamSearch known answers (recalled constants): decoded paths exact,
    scores to the printed digits; the beam test needs top_paths = 2 (both paths and both scores)."""
    g = _tf_upstream()
    gr = g["greedy"]
    out, n, sc = O.greedy(np.asarray(gr["input_prob_matrix"], np.float32), seq_len=gr["seq_len"], eps=1e-30)
    for b, want in enumerate(gr["decoded"]):
        assert out[b, :n[b]].tolist() == want and np.all(out[b, n[b]:] == -1)
    np.testing.assert_allclose(sc, [np.sum(-np.log(f)) for f in gr["neg_log_prob_factors"]], rtol=1e-6)
    bm = g["beam"]
    probs = np.asarray(bm["input_prob_matrix"], np.float32)[None]
    out, n, lp = O.beam_topk(probs, bm["top_paths"], seq_len=[bm["seq_len"]], beam_width=bm["beam_width"], merge_repeated=bm["merge_repeated"], eps=0.0)
    for pth, want in enumerate(bm["decoded"]):
        assert out[0, pth, :n[0, pth]].tolist() == want
    np.testing.assert_allclose(lp[0], bm["log_prob"], rtol=0, atol=1.5e-6)
    # top-1 entry point agrees with path 0 of the top-k one
    o1, n1, l1 = O.beam(probs, seq_len=[bm["seq_len"]], beam_width=bm["beam_width"], merge_repeated=bm["merge_repeated"], eps=0.0)
    assert o1[0, :n1[0]].tolist() == bm["decoded"][0] and abs(l1[0] - lp[0, 0]) == 0
