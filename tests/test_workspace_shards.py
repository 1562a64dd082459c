"""CPU: shard bookkeeping of the per-rank workspace (landiff_b200/dit.py `_workspace`): for every ring size the
shards' text / image row counts and first-image-token offsets tile the 17 776-token sequence exactly, K|V live in one
contiguous buffer (a ring hop is a single copy), and the main net keeps an fp32 residual stream while the control net
stays bf16.  Only buffer shapes are touched — nothing is computed, so no GPU is needed."""
import pytest
import torch

from landiff_b200 import parallel
from landiff_b200.factory import FULL, TINY, build_warp


@pytest.fixture(scope="module")
def nets():
    # TINY widths keep the buffers small; the token geometry is what is under test, so borrow the full-shape frame grid
    warp = build_warp(TINY)
    return warp.control_model.diffusion_model, warp.main_model.diffusion_model


@pytest.mark.parametrize("sp", [1, 2, 4])
def test_shards_tile_the_sequence(nets, sp):
    ctrl, main = nets
    T, H, W = 13, 20, 22           # 13 * 10 * 11 = 1430 image tokens + 6 text tokens (TINY) = 1436 = 4 * 359
    n_total = main.text_length + T * (H // 2) * (W // 2)
    if n_total % sp:
        pytest.skip("sequence does not split evenly")
    seen_txt = seen_img = 0
    next_g0 = 0
    for r in range(sp):
        lay = parallel.Layout(2 * sp, r, 2, sp)
        for net in (ctrl, main):
            net.sp_layout, net._ws = lay, {}
        ws = main._workspace(1, T, H, W, "cpu")
        start, count = parallel.shard_bounds(n_total, sp, r)
        assert (ws["shard"].start, ws["shard"].count, ws["n_total"]) == (start, count, n_total)
        assert ws["n_txt"] + ws["n_img"] == count
        assert ws["n_txt"] == max(min(start + count, main.text_length) - start, 0)
        if ws["n_img"]:
            assert ws["g0"] == next_g0
            next_g0 += ws["n_img"]
        seen_txt += ws["n_txt"]
        seen_img += ws["n_img"]
        # buffers: rows = shard size; K|V contiguous; head-major q / k / v
        assert ws["hidden"].shape == (1, count, main.hidden_size) and ws["hidden"].dtype == torch.float32
        assert ws["q"].shape == (1, main.num_attention_heads, count, 64)
        assert ws["kv"].is_contiguous() and ws["k"].data_ptr() == ws["kv"].data_ptr()
        assert ws["v"].data_ptr() == ws["kv"].data_ptr() + ws["k"].numel() * 2
        wc = ctrl._workspace(1, T, H, W, "cpu")
        assert wc["hidden"].dtype == torch.bfloat16 and wc["ctrl"].shape == (ctrl.num_layers, 1, count, ctrl.hidden_size)
        assert main._workspace(1, T, H, W, "cpu") is ws      # cached: a warmed-up step allocates nothing
    assert seen_txt == main.text_length and seen_img == n_total - main.text_length
    for net in (ctrl, main):
        net.sp_layout, net._ws = None, {}


def test_full_shape_token_count_splits_for_every_ring_size():
    assert FULL.n_tok == 17776 == 16 * 1111
    for sp in (2, 4, 8):
        assert FULL.n_tok % sp == 0
