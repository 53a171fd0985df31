"""A timing MODEL of the matrix peer exchange (csrc/mimo.cu: PeerPub / peer_store / peer_signal, K3's flag wait in
csrc/fft_kernels.cuh, and the overlapped finish of fcb_mimo_set_overlap) — not the CUDA code itself: the GPU tests and
bench.py's `mimo` block check that.  What this pins is the ORDER the protocol rests on, under arbitrary kernel durations
and NVLink delays, for 2 .. 8 shards:

  * inbox slots and flags are double-buffered by block parity, so the reduce of block n + 2 on ANY shard overwrites what
    K3 of block n on every other shard reads; nothing but "a shard's reduce comes after its own previous K3" (stream order,
    or the ev_fin wait in front of the reduce when K3 has its own stream) keeps it from starting too early;
  * K3 waits for flag == n exactly, so a flag that jumps from n - 2 to n + 2 past a late K3 would hang it.

The model computes start / end times from the dependencies as the host code enqueues them and asserts that for every
block n and every pair of shards, reduce(n + 2) on one starts only after K3(n) on the other has ended.  The negative
control removes the one wait and must find the overwrite.
"""
import random

import pytest


def simulate(G, nblocks, rng, overlap, keep_order=True):
    """returns (times, violations): kernels' (start, end) per shard and block, and the (n, writer, reader) triples where
    reduce(n + 2) of `writer` starts before K3(n) of `reader` has ended"""
    mac = [[rng.uniform(5, 90) for _ in range(nblocks + 1)] for _ in range(G)]
    red = [[rng.uniform(1, 12) for _ in range(nblocks + 1)] for _ in range(G)]
    k3 = [[rng.uniform(1, 40) for _ in range(nblocks + 1)] for _ in range(G)]
    link = [[[rng.uniform(0.5, 30) for _ in range(nblocks + 1)] for _ in range(G)] for _ in range(G)]  # g -> r, block n
    red_t = [[(0.0, 0.0)] * (nblocks + 1) for _ in range(G)]
    k3_t = [[(0.0, 0.0)] * (nblocks + 1) for _ in range(G)]
    main_free = [0.0] * G  # when the shard's main stream has run everything enqueued so far
    for n in range(1, nblocks + 1):
        for r in range(G):
            # partial_dev: K1 || MAC, then the reduce (which stores into every inbox and raises the flags at its end)
            mac_end = main_free[r] + mac[r][n]
            start = mac_end
            if overlap and keep_order:
                start = max(start, k3_t[r][n - 1][1])  # cudaStreamWaitEvent(stream, ev_fin) in front of the reduce
            red_t[r][n] = (start, start + red[r][n])
            main_free[r] = red_t[r][n][1]
        for r in range(G):
            # finish_dev: K3 after its own block's reduce (stream order, or ev_red), spinning until every flag says n
            ready = max(red_t[g][n][1] + link[g][r][n] for g in range(G))
            if overlap:
                start = max(red_t[r][n][1], k3_t[r][n - 1][1])  # the finish stream runs K3s in order
            else:
                start = main_free[r]
            begin = max(start, ready)
            k3_t[r][n] = (start, begin + k3[r][n])
            if not overlap:
                main_free[r] = k3_t[r][n][1]
    bad = []
    for n in range(1, nblocks - 1):
        for w in range(G):
            for r in range(G):
                if red_t[w][n + 2][0] < k3_t[r][n][1]:
                    bad.append((n, w, r))
    return (red_t, k3_t), bad


@pytest.mark.parametrize("G", [2, 3, 4, 8])
@pytest.mark.parametrize("overlap", [False, True])
def test_no_shard_overwrites_a_slot_still_being_read(G, overlap):
    for seed in range(200):
        _, bad = simulate(G, 40, random.Random(1000 * G + seed), overlap)
        assert not bad, (seed, bad[:3])


def test_the_model_sees_the_overwrite_when_the_wait_is_removed():
    """negative control: K3 on its own stream WITHOUT the ev_fin wait in front of the next reduce"""
    hits = sum(bool(simulate(4, 40, random.Random(seed), True, keep_order=False)[1]) for seed in range(50))
    assert hits >= 40


def test_overlap_shortens_the_cycle_in_the_model():
    """the point of the overlapped finish: per-block time tends to max(MAC + reduce, exchange + K3), not their sum"""
    rng_a, rng_b = random.Random(7), random.Random(7)
    (_, k3_seq), _ = simulate(4, 200, rng_a, False)
    (_, k3_ovl), _ = simulate(4, 200, rng_b, True)
    assert max(t[200][1] for t in k3_ovl) < 0.9 * max(t[200][1] for t in k3_seq)
