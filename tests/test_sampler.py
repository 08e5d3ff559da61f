"""Sampler row: the restatement of reference sampler.c against the unmodified reference
(when it is mounted) and against the committed fixture made from it; on the GPU, the device-side sampler against
the restatement on the same logits and generator state."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT  # noqa: E402

from calm_b200 import modelgen as mg  # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "sampler.npz")
CASES = [(1.0, 0.1), (0.7, 0.05), (1.5, 0.3), (0.0, 0.1), (1.0, 1.0)]  # (temperature, minp); the last two are greedy


def fixture_logits():
    rng = np.random.default_rng(11)
    flat = rng.standard_normal((8, 1500)).astype(np.float32) * 0.3   # nearly every token survives
    peaked = rng.standard_normal((8, 1500)).astype(np.float32) * 4.0  # a handful survive
    return np.concatenate([flat, peaked])


def run_sampler(S, kind):
    out = []
    for t, m in CASES:
        s = S(kind, t, m, 0x9E3779B97F4A7C15)
        toks = [s.sample(l) for l in fixture_logits() for _ in range(4)]
        out.append((toks, s.rng_state))
    return out


def test_restatement_matches_fixture_made_from_the_reference(oracle_pkg):
    fx = np.load(FIX)
    for i, (toks, rng) in enumerate(run_sampler(oracle_pkg.Sampler, "port")):
        assert rng == int(fx["rng"][i])  # integer generator: bit exact
        same = np.mean(np.array(toks) == fx["tokens"][i])
        # the float path may differ from the reference's -ffast-math build by an ulp in a bin edge: at most a stray sample
        assert same >= 0.98, (CASES[i], same)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcalm_ref_sampler.so")), reason="reference sampler not built")
def test_fixture_is_the_live_reference(oracle_pkg):
    fx = np.load(FIX)
    for i, (toks, rng) in enumerate(run_sampler(oracle_pkg.Sampler, "reference")):
        assert rng == int(fx["rng"][i])
        assert np.array_equal(np.array(toks), fx["tokens"][i])


@pytest.mark.gpu
@pytest.mark.parametrize("temperature,minp", [(1.0, 0.1), (0.8, 0.02), (1.3, 0.5)])
def test_device_sampler_follows_the_restatement(oracle_pkg, temperature, minp):
    from calm_b200 import lib

    spec = mg.SPECS["tiny-llama"]
    model = mg.HostModel(spec, seed=0)
    dm = lib.DeviceModel(spec, model.tensors)
    try:
        ref = oracle_pkg.Sampler("port", temperature, minp, 1234567)
        rng, tok, agree, n = 1234567, 5, 0, 48
        for pos in range(n):
            toks, rng = dm.decode_sample(tok, pos, 1, temperature, minp, rng)
            want = ref.sample(dm.device_logits())   # same logits, same generator state
            assert rng == ref.rng_state             # xorshift*: bit exact
            agree += int(toks[0] == want)
            tok = int(toks[0])
        assert agree >= n - 1, (agree, n)            # device expf vs libm: at most a stray bin edge
        # the device-resident loop draws the same sequence as stepping it from the host
        a, ra = dm.decode_sample(5, 0, 16, temperature, minp, 99)
        seq, r, t = [], 99, 5
        for pos in range(16):
            o, r = dm.decode_sample(t, pos, 1, temperature, minp, r)
            seq.append(int(o[0]))
            t = int(o[0])
        assert list(a) == seq and ra == r
        # greedy settings fall back to the arg-max loop and leave the generator alone
        g, rg = dm.decode_sample(5, 0, 8, 0.0, minp, 77)
        assert rg == 77 and list(g) == list(dm.decode_greedy(5, 0, 8))
    finally:
        dm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("vocab,scale,temperature,minp", [(128256, 0.3, 1.0, 0.1), (128256, 0.3, 1.0, 0.02), (128256, 4.0, 0.8, 0.05), (128256, 0.05, 1.0, 0.3),
                                                          (151936, 1.0, 1.2, 0.01), (32000, 0.5, 1.0, 0.1), (1000, 0.3, 1.0, 0.1)])
def test_device_sampler_at_real_vocabulary(oracle_pkg, vocab, scale, temperature, minp):
    """The sampler KERNELS (k_sample_scan + k_sample_pick) on Llama-3 / Qwen2 / Mistral sized vocabularies: flat logits
    (tens of thousands of survivors -> the chunk-sum walk), peaked logits (a handful), many chunks and a ragged tail,
    against the oracle's restatement of sample() with the same generator state."""
    from calm_b200 import lib

    rng = np.random.default_rng(vocab + int(scale * 100))
    ref = oracle_pkg.Sampler("port", temperature, minp, 4242)
    state, agree, n, surv = 4242, 0, 24, []
    us_max, worst = 0.0, 0
    for i in range(n):
        logits = (rng.standard_normal(vocab) * scale).astype(np.float32)
        surv.append(int((logits >= logits.max() + np.log(minp) * temperature).sum()))
        tok, state, us = lib.sample_logits(logits, temperature, minp, state, timed=True)
        us_max = max(us_max, us)
        want = ref.sample(logits)
        assert state == ref.rng_state  # xorshift*: bit exact
        agree += int(tok == want)
        # whatever the rounding of the sums, the pick must be a survivor and sit next to the reference's pick in survivor order
        cutoff = logits.max() + np.float32(np.log(np.float32(minp))) * np.float32(temperature)
        keep = np.nonzero(logits >= cutoff)[0]
        assert tok in keep
        dist = abs(int(np.searchsorted(keep, tok)) - int(np.searchsorted(keep, want)))
        worst = max(worst, dist)
    print(f"vocab {vocab} scale {scale} T {temperature} minp {minp}: survivors {min(surv)}..{max(surv)}, agree {agree}/{n}, worst rank distance {worst}, sampler kernels <= {us_max:.1f} us")
    if max(surv) <= 2048:  # the reference's own additions in the reference's order: at most a stray expf-vs-libm bin edge
        assert agree >= n - 1, (agree, n, max(surv))
    else:
        # tens of thousands of survivors: a bin is ~1e-5 of the sum, less than the difference between a device expf and libm summed
        # over the vocabulary, so equality is not defined; the pick stays within a few survivors of the reference's
        assert worst <= max(4, max(surv) // 4000), (worst, max(surv))


def _device_algorithm(logits, temperature, minp, rng_state, chunk=1024, exact=2048):
    """Python mirror of k_sample_scan + k_sample_pick (stages.cuh): chunked compaction in index order, then the exact walk
    (<= `exact` survivors) or the chunk-sum walk.  float32 arithmetic step by step, like the one device thread."""
    f = np.float32
    s = rng_state
    s ^= s >> 12
    s ^= (s << 25) & 0xFFFFFFFFFFFFFFFF
    s ^= s >> 27
    coin = f(((s * 0x2545F4914F6CDD1D & 0xFFFFFFFFFFFFFFFF) >> 32) >> 8) / f(16777216.0)
    mx = f(logits.max())
    cutoff = f(mx + f(f(np.log(f(minp))) * f(temperature)))
    chunks = []
    for c0 in range(0, len(logits), chunk):
        l = logits[c0:c0 + chunk]
        keep = np.nonzero(l >= cutoff)[0]
        probs = np.exp(((l[keep] - mx) / f(temperature)).astype(f)).astype(f)
        csum = f(0)
        for p in probs:
            csum = f(csum + p)
        chunks.append((keep + c0, probs, csum))
    total = sum(len(k) for k, _, _ in chunks)
    fallback = max((int(k[-1]) for k, _, _ in chunks if len(k)), default=0)
    cum = f(0)
    if total <= exact:
        for _, probs, _ in chunks:
            for p in probs:
                cum = f(cum + p)
    else:
        for _, _, cs in chunks:
            cum = f(cum + cs)
    r = f(coin * cum)
    cdf = f(0)
    for idx, probs, cs in chunks:
        if total > exact and not (r < f(cdf + cs)):
            cdf = f(cdf + cs)
            continue
        for i, p in zip(idx, probs):
            cdf = f(cdf + p)
            if r < cdf:
                return int(i), s
    return fallback, s


@pytest.mark.parametrize("vocab,scale", [(1000, 0.3), (4128, 0.3), (4128, 4.0), (9000, 0.2)])
def test_device_sampling_algorithm_agrees_with_the_restatement(oracle_pkg, vocab, scale):
    """The chunked algorithm the device kernels implement (several chunks, a ragged last chunk, the chunk-sum path beyond
    2048 survivors) against oracle_sample on the same logits: identical draws up to a stray bin edge."""
    rng = np.random.default_rng(vocab)
    agree = n = 0
    for t, m in [(1.0, 0.1), (0.7, 0.02)]:
        ref = oracle_pkg.Sampler("port", t, m, 42)
        state = 42
        for _ in range(40):
            logits = (rng.standard_normal(vocab) * scale).astype(np.float32)
            tok, state = _device_algorithm(logits, t, m, state)
            want = ref.sample(logits)
            assert state == ref.rng_state
            agree += int(tok == want)
            n += 1
    assert agree >= n - 1, (agree, n)
