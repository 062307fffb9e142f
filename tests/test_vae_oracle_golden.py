"""Row V oracle (oracle/vae_oracle.py) against fixtures produced by the reference's own AutoencoderKLCogVideoX
(tests/golden/make_golden.py:gen_vae).  CPU only."""
import os

import pytest
import torch

from oracle import vae_oracle as V


@pytest.fixture(scope="module")
def fix(golden_dir):
    return torch.load(os.path.join(golden_dir, "vae_tiny.pt"), weights_only=False)


@pytest.fixture(scope="module")
def setup(fix):
    cfg = V.VaeConfig(**fix["cfg"])
    p = V.synth_decoder_params(cfg, seed=fix["seed"])
    assert abs(float(sum(v.double().sum() for v in p.values())) - fix["weight_checksum"]) < 1e-6   # RNG canary
    return cfg, p


def _close(a, b, tol=2e-5):
    assert a.shape == b.shape
    err = float((a - b).abs().max() / b.abs().max())
    assert err < tol, err


def test_frame_batches_and_tile_plan():
    assert V.frame_batches(13, 2) == [(0, 3), (3, 5), (5, 7), (7, 9), (9, 11), (11, 13)]     # 6 decoder calls
    assert V.frame_batches(1, 2) == [(0, 3)]   # the end index overshoots like the reference's; slicing clamps it
    assert V.frame_batches(2, 2) == [(0, 2)] and V.frame_batches(5, 2) == [(0, 3), (3, 5)]
    tp = V.tile_plan(V.VaeConfig(), 60, 90)                                                  # the shipped 480x720 geometry
    assert tp["rows"] == [0, 25, 50] and tp["cols"] == [0, 36, 72]                           # 9 tiles
    assert (tp["blend_h"], tp["blend_w"], tp["limit_h"], tp["limit_w"]) == (40, 72, 200, 288)


def test_decoder_call_chain_with_conv_cache(fix, setup):
    cfg, p = setup
    zt = fix["z"][:, :, :, :4, :6]
    with torch.no_grad():
        y0, cache = V.decoder_forward(p, cfg, zt[:, :, :3], None)
        y1, _ = V.decoder_forward(p, cfg, zt[:, :, 3:5], cache)
    _close(y0, fix["decoder_chain"]["y0"])
    _close(y1, fix["decoder_chain"]["y1"])


def test_decode_untiled(fix, setup):
    cfg, p = setup
    with torch.no_grad():
        _close(V.decode(p, cfg, fix["z"], use_tiling=False), fix["decode_untiled"])


def test_decode_tiled_blended(fix, setup):
    cfg, p = setup
    with torch.no_grad():
        out = V.decode(p, cfg, fix["z"], use_tiling=True)
        _close(out, fix["decode_tiled"])
        z2 = torch.cat([fix["z"], 0.5 * fix["z"].flip(3)], dim=0)
        s = float(V.decode(p, cfg, z2).double().sum())
    assert abs(s - fix["decode_tiled_b2_sum"]) < 1e-3 * abs(fix["decode_tiled_b2_sum"]) + 1e-2
