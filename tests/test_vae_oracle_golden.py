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


# ------------------------------------------------------------------------------------------- host glue
def _post_fix(golden_dir):
    import numpy as np
    d = np.load(os.path.join(golden_dir, "postprocess.npz"))
    video = torch.from_numpy(d["video_bf16_bits"].copy()).view(torch.bfloat16)
    return d, video


def test_frames_uint8_oracle_matches_the_reference_video_processor(golden_dir):
    """oracle.frames_uint8 == the reference's VideoProcessor.postprocess_video + export_to_video / numpy_to_pil conversions,
    bit for bit (fixture: tests/golden/make_golden.py:gen_post)."""
    import numpy as np
    d, video = _post_fix(golden_dir)
    assert np.array_equal(V.frames_uint8(video, round_half_even=False), d["trunc"])
    assert np.array_equal(V.frames_uint8(video, round_half_even=True), d["rounded"])
    assert (d["trunc"] != d["rounded"]).any()   # the two roundings are distinguishable on this fixture


def test_postprocess_video_np_matches_reference_and_export_to_video_writes_mp4(golden_dir, tmp_path):
    """The host surface: postprocess_video('np') reproduces the reference's fp32 frames; export_to_video accepts them (and
    the uint8 frames) and writes a readable mp4 with the frame count and size of the input (D/utils/export_utils.py:116-186)."""
    import numpy as np
    import s2v_b200
    d, video = _post_fix(golden_dir)
    got = s2v_b200.postprocess_video(video, "np")
    assert got.dtype == np.float32 and np.array_equal(got, d["as_np"])
    with pytest.raises(RuntimeError):
        s2v_b200.postprocess_video(video, "uint8")          # host tensor: no CPU substitute for the device conversion
    cv2 = pytest.importorskip("cv2")
    big = np.kron(d["trunc"][0], np.ones((1, 8, 8, 1), dtype=np.uint8))      # 64x96 frames (codecs dislike 8x12)
    for frames in (list(big), [f.astype(np.float32) / 255.0 for f in big]):
        path = s2v_b200.export_to_video(frames, str(tmp_path / "o.mp4"), fps=8)
        cap = cv2.VideoCapture(path)
        n = 0
        while True:
            ok, img = cap.read()
            if not ok:
                break
            assert img.shape == (64, 96, 3)
            n += 1
        assert n == len(frames)


# ------------------------------------------------------------------------------------------- encoder
@pytest.fixture(scope="module")
def enc_fix(golden_dir):
    return torch.load(os.path.join(golden_dir, "vae_enc_tiny.pt"), weights_only=False)


def _image_tensor(img):
    import numpy as np
    x = torch.from_numpy(np.expand_dims(img, 0)).float() / 255.0 * 2.0 - 1.0
    return x.permute(0, 3, 1, 2).unsqueeze(0).permute(0, 2, 1, 3, 4)


def test_encoder_oracle_matches_reference_encode(enc_fix):
    """oracle encoder (single-frame path) == the reference AutoencoderKLCogVideoX.encode moments: one tile-sized call, the
    untiled image, and the tiled + blended image (fixture: tests/golden/make_golden.py:gen_vae_enc)."""
    cfg = V.VaeConfig(**enc_fix["cfg"])
    p = V.synth_encoder_params(cfg, seed=enc_fix["seed"])
    assert abs(float(sum(v.double().sum() for v in p.values())) - enc_fix["weight_checksum"]) < 1e-6   # RNG canary
    x = _image_tensor(enc_fix["image"])
    with torch.no_grad():
        _close(V.encoder_forward(p, cfg, x[:, :, :, :32, :48]), enc_fix["moments_tile"])
        _close(V.encode_moments(p, cfg, x, use_tiling=False), enc_fix["moments_untiled"])
        _close(V.encode_moments(p, cfg, x, use_tiling=True), enc_fix["moments_tiled"])
        assert enc_fix["moments_tiled"].shape[3] == 9    # the tiled result has the reference's own (odd) geometry at this size
        ref = V.reference_image_latents(p, cfg, enc_fix["image"], enc_fix["noise"], use_tiling=True)
        _close(ref, enc_fix["ref_img_states"])
