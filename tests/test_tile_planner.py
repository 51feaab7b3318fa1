"""Host logic of the tensor-core conv kernel: the tile planner (k2y_tc_plan) — no GPU needed (assumes a B200).

Pins the choices the device timelines in profiles/r01_tile_model.md led to, and the invariants the kernel relies on."""
import ctypes

import pytest

from k210_yolo_framework_b200._lib import MATH_TC_3XTF32, MATH_TC_BF16X3, MATH_TC_TF32, K2YError, check, lib


def plan(M, N, K, ksize, mode=MATH_TC_BF16X3):
    v = [ctypes.c_int() for _ in range(5)]
    check(lib.k2y_tc_plan(M, N, K, ksize, mode, *[ctypes.byref(x) for x in v]))
    return dict(zip(("bn", "splits", "stages", "cluster", "math"), (x.value for x in v)))


def test_cfg2_layer_choices():
    # head 3x3 convs: few m-tiles, very deep K -> wide 2-stage tiles + split-K (profiles/r01_tile_model.md)
    assert plan(2240, 192, 6912, 3) == {"bn": 192, "splits": 8, "stages": 2, "cluster": 2, "math": MATH_TC_BF16X3}
    assert plan(8960, 128, 4608, 3) == {"bn": 128, "splits": 2, "stages": 2, "cluster": 2, "math": MATH_TC_BF16X3}
    # conv_pw_1: K = 24 fits one 32-wide tf32 k-block -> 3xTF32, one n-tile, no cluster; its pixel-pair form is a K=48, N=96 GEMM
    p = plan(573440, 48, 24, 1)
    assert (p["bn"], p["splits"], p["cluster"], p["math"]) == (48, 1, 1, MATH_TC_3XTF32)
    p = plan(286720, 96, 48, 1)
    assert (p["bn"], p["splits"], p["cluster"], p["math"]) == (96, 1, 1, MATH_TC_BF16X3)
    # K-deep pointwise layers on 70 m-tiles: CTA pairs (weight multicast), no split
    p = plan(8960, 384, 384, 1)
    assert p["cluster"] == 2 and p["splits"] == 1 and p["bn"] in (96, 192)


@pytest.mark.parametrize("mode", [MATH_TC_3XTF32, MATH_TC_TF32, MATH_TC_BF16X3])
def test_planner_invariants(mode):
    shapes = [(M, N, K, ks) for M in (128, 2240, 8960, 35840, 573440, 1478656) for N in (16, 48, 75, 96, 128, 192, 255, 384, 768, 1024)
              for K, ks in ((16, 1), (24, 1), (96, 1), (384, 1), (768, 1), (288, 3), (1152, 3), (4608, 3), (6912, 3))]
    for M, N, K, ks in shapes:
        p = plan(M, N, K, ks, mode)
        n16 = (N + 15) // 16 * 16
        assert 16 <= p["bn"] <= min(256, n16) and p["bn"] % 16 == 0
        n_tiles = -(-n16 // p["bn"])
        assert n_tiles == 1 or p["bn"] % 32 == 0                      # TMA-store boxes are 32 columns wide
        assert 2 <= p["stages"] <= 8
        three_x = p["math"] != MATH_TC_TF32
        if three_x:
            assert 2 * p["bn"] + 64 * p["stages"] <= 512              # TMEM: two accumulators + A planes of every stage
        a_bytes = 32768 if p["math"] == MATH_TC_BF16X3 else 16384
        assert p["stages"] * (a_bytes + p["bn"] * 128 * (2 if three_x else 1)) + 43400 <= 232448   # shared memory: stage ring + barriers + epilogue staging (FIXED_SMEM)
        nkb = -(-K // (64 if p["math"] == MATH_TC_BF16X3 else 32))
        assert 1 <= p["splits"] <= 8 and (p["splits"] == 1 or (N % 4 == 0 and nkb // p["splits"] >= 6))
        assert p["cluster"] in (1, 2)
        if mode != MATH_TC_BF16X3:
            assert p["math"] == mode
        elif ks == 1:
            assert p["math"] == (MATH_TC_3XTF32 if K <= 32 else MATH_TC_BF16X3)
        else:
            # 3x3: whole 64-channel k-blocks per tap, or 32 channels (two taps per k-block); anything else runs as 3xTF32
            assert p["math"] == (MATH_TC_BF16X3 if ((K // 9) % 64 == 0 or K // 9 == 32) else MATH_TC_3XTF32)


def test_planner_rejects_bad_arguments():
    with pytest.raises(K2YError):
        plan(0, 48, 24, 1)
    with pytest.raises(K2YError):
        plan(128, 48, 24, 2)
    with pytest.raises(K2YError):
        plan(128, 48, 24, 1, mode=0)    # the fp32 CUDA-core mode has no tiles to plan


def plan_budget(M, N, K, ksize, sm_limit, mode=MATH_TC_BF16X3):
    v = [ctypes.c_int() for _ in range(6)]
    check(lib.k2y_tc_plan_budget(M, N, K, ksize, mode, sm_limit, *[ctypes.byref(x) for x in v]))
    return dict(zip(("bn", "splits", "stages", "cluster", "math", "grid"), (x.value for x in v)))


# every tensor-core GEMM of cfg 2 (yolo_mobilev1-0.75 at batch 32): (M, N, K, ksize)
CFG2_GEMMS = [(286720, 96, 48, 1), (143360, 96, 48, 1), (143360, 96, 96, 1), (35840, 192, 96, 1), (35840, 192, 192, 1), (8960, 384, 192, 1),
              (8960, 384, 384, 1), (2240, 768, 384, 1), (2240, 768, 768, 1), (2240, 192, 6912, 3), (2240, 75, 192, 1), (2240, 128, 768, 1),
              (8960, 128, 4608, 3), (8960, 75, 128, 1)]


@pytest.mark.parametrize("sm_limit", [0, 74, 72, 48, 9])
def test_planner_under_sm_budget(sm_limit):
    """k2y_net_set_sm_limit: the persistent grid never exceeds the budget (rounded down to even, at least 8), and every layer of
    the headline network still gets a tile that fits shared memory and tensor memory."""
    budget = 148 if sm_limit in (0,) or sm_limit >= 148 else max(8, sm_limit) & ~1
    for M, N, K, ks in CFG2_GEMMS:
        p = plan_budget(M, N, K, ks, sm_limit)
        assert 1 <= p["grid"] <= budget and p["grid"] % p["cluster"] == 0
        assert 2 <= p["stages"] <= 8 and 16 <= p["bn"] <= 256
        three_x = p["math"] != MATH_TC_TF32
        if three_x:
            assert 2 * p["bn"] + 64 * p["stages"] <= 512
        a_bytes = 32768 if p["math"] == MATH_TC_BF16X3 else 16384
        assert p["stages"] * (a_bytes + p["bn"] * 128 * (2 if three_x else 1)) + 43400 <= 232448
        if sm_limit == 0:
            q = plan(M, N, K, ks)
            assert {k: p[k] for k in q} == q        # budget 0 = the plain planner


def test_budget_uses_its_whole_share_on_the_small_grid_layers():
    # conv_pw_7..11 (70 m-tiles x n-tiles): on the whole device one CTA per SM, on a 74-SM share 74 CTAs with more tiles each
    assert plan_budget(8960, 384, 384, 1, 0)["grid"] > 100
    assert plan_budget(8960, 384, 384, 1, 74)["grid"] == 74
    with pytest.raises(K2YError):
        plan_budget(8960, 384, 384, 1, -1)
