"""Host-side model of why the attention grid is (heads, query-tile pairs) with the heads fastest (leopard_b200/csrc/lp_attn3.cu,
lp_attn_bwd.cu): the block scheduler hands out CTAs in linear block order to whichever SM frees up first (list scheduling), and a
causal pair qp costs 2 qp + 2 key-tile steps.  With the pairs fastest only each head's own pairs were ordered heaviest-first; the
last heads' heaviest CTAs started near the end of the kernel and ran alone.  Measured on the B200: 0.208 -> 0.174 ms at S = 4796
(905 -> 1082 TFLOP/s), 707 -> 876 TFLOP/s at S = 2937 (profiles/r02y_attn_time_lpt.log)."""
import heapq


def makespan(costs, n_sm=148):
    free = [0] * n_sm
    heapq.heapify(free)
    end = 0
    for c in costs:
        t = heapq.heappop(free) + c
        end = max(end, t)
        heapq.heappush(free, t)
    return end


def pair_cost(qp, T, step=3960, fixed=5000):      # cycles per 128-key step of a CTA and per-CTA fixed cost, from the phase trace
    n_total = -(-T // 128)
    return min(n_total, 2 * qp + 2) * step + fixed


def orders(T, heads=32):
    n_pairs = -(-T // 256)
    pairs_fastest = [pair_cost(n_pairs - 1 - x, T) for _ in range(heads) for x in range(n_pairs)]      # grid (pairs, heads): the old order
    heads_fastest = [pair_cost(qp, T) for qp in range(n_pairs - 1, -1, -1) for _ in range(heads)]      # grid (heads, pairs): now
    return pairs_fastest, heads_fastest


def test_heaviest_first_over_the_whole_grid_removes_the_tail():
    for T, min_gain in ((4796, 1.20), (2937, 1.20), (16384, 1.04)):
        old, new = orders(T)
        ideal = sum(new) / 148
        assert makespan(new) <= 1.09 * ideal, T           # within a few per cent of perfect balance (the heaviest CTA bounds short sequences)
        assert makespan(old) >= min_gain * makespan(new), T


def test_backward_dkv_grid_has_the_same_problem_and_fix():
    # dK/dV kernel: 8 kv heads x 32 kv tiles at seq 4096; tile k sees 32 - k query tiles
    cost = lambda k: (32 - k) * 4000 + 5000
    old = [cost(k) for _ in range(8) for k in range(32)]
    new = [cost(k) for k in range(32) for _ in range(8)]
    assert makespan(new) == max(new)                      # bounded by the single heaviest CTA: cannot do better without splitting it
    assert makespan(old) > 1.15 * makespan(new)
