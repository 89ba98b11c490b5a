"""-m gpu: fp16 / bf16 network outputs (what the reference decodes under --amp,
src/sdnet/model/trainer.py:40-42,142-157).  ATen rounds the sigmoid and the clamp to the tensor
dtype, so scores live on a 10- / 7-bit mantissa grid and ties are everywhere; the kernels must
reproduce that exactly."""
import numpy as np
import pytest
import torch

from oracle import sdnet_oracle as O
from oracle import torch_port as TP
from structuredetector_b200 import Decoder, ops
from structuredetector_b200.synth import CONFIGS, DecodeConfig, make_raw, split_outputs
from tests.helpers import assert_packed_equal, make_args, packed_np, plain

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]
# margins of Num<DT> in csrc/sdnet_decode.cu
MARGINS = {torch.float16: (0.02, 3.0, -11.0, 0.15, 5.0), torch.bfloat16: (0.1, 2.0, -13.0, 0.6, 4.0)}


def _all_finite_values(dtype, device):
    bits = torch.arange(0, 1 << 16, dtype=torch.int32, device=device).to(torch.int16)
    x = bits.view(dtype)
    return x[torch.isfinite(x)]


def _device_activation(dtype, device):
    """float32 ndarray -> the reference's clamped sigmoid evaluated in `dtype` on the device."""
    def fn(m):
        t = torch.from_numpy(np.ascontiguousarray(m)).to(device).to(dtype)
        return torch.clamp(torch.sigmoid(t), min=1e-6, max=1 - 1e-6).float().cpu().numpy()
    return fn


@pytest.mark.parametrize("dtype", DTYPES)
def test_activation_bit_exact_for_every_value(cuda_device, dtype):
    x = _all_finite_values(dtype, cuda_device)
    pad = (-x.numel()) % 64
    xp = torch.cat([x, x[:pad]]).view(1, 1, -1, 64)
    got = ops.activate_maps(xp)
    want = torch.clamp(torch.sigmoid(xp), min=1e-6, max=1 - 1e-6)
    assert got.dtype == dtype and torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("dtype", DTYPES)
def test_score_function_monotone_and_margins_hold(cuda_device, dtype):
    """Exhaustive over all 65,536 inputs: S_T is monotone, and outside the near-tie margins the
    kernel assumes (Num<DT>::kNear/kHi/kLo, kSatX) two logits never share a score."""
    near, hi, lo, near2, hi2 = MARGINS[dtype]
    x = torch.sort(_all_finite_values(dtype, cuda_device).float()).values
    x = x[(x >= -30) & (x <= 30)]
    pad = (-x.numel()) % 64
    s = ops.activate_maps(torch.cat([x, x[:pad]]).to(dtype).view(1, 1, -1, 64)).float().view(-1)[: x.numel()]
    assert bool((s[1:] >= s[:-1]).all())
    # for every h in [lo, hi]: the largest representable x with x < h - near scores strictly lower
    xs, ss = x.cpu().numpy().astype(np.float64), s.cpu().numpy()
    mid = np.flatnonzero((xs >= lo) & (xs <= hi))
    j = np.searchsorted(xs, xs[mid] - near, side="left") - 1  # last index with xs[j] < h - near
    ok = j >= 0
    assert (ss[j[ok]] < ss[mid[ok]]).all()
    # second zone: h in (hi, hi2] with the wider margin
    mid2 = np.flatnonzero((xs > hi) & (xs <= hi2))
    j2 = np.searchsorted(xs, xs[mid2] - near2, side="left") - 1
    assert (ss[j2] < ss[mid2]).all()
    # beyond it: x <= hi2 - 1 never ties with h > hi2
    assert ss[np.searchsorted(xs, hi2 - 1.0, side="right") - 1] < ss[np.searchsorted(xs, hi2, side="right")]
    # saturation: every |x| >= 14 scores like +-14
    top, bot = ss[np.searchsorted(xs, 14.0)], ss[np.searchsorted(xs, -14.0, side="right") - 1]
    assert (ss[xs >= 14.0] == top).all() and (ss[xs <= -14.0] == bot).all()


CASES = [("cfg1", "noise", 3), ("cfg1", "blobs", 2), ("cfg1", "ties", 2), ("cfg3", "noise", 2), ("cfg3", "blobs", 1),
         ("cfg4", "noise", 1)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,mode,batch", CASES)
def test_decode_matches_oracle(cuda_device, dtype, name, mode, batch):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=batch).to(dtype)  # the network's reduced-precision output
    outs = split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts)
    got = packed_np(ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh))
    f32 = raw.float().numpy()
    m, n = cfg.labels, cfg.parts
    want = O.decode_packed(f32[:, :m], f32[:, m:m + n], f32[:, m + n:m + n + 2], f32[:, m + n + 2:], cfg.max_objects,
                           cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh,
                           activation_fn=_device_activation(dtype, cuda_device),
                           conf_cmp=float(torch.tensor(cfg.conf_threshold, dtype=dtype)))
    assert_packed_equal(got, want, what=f"{dtype} {name}/{mode}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("h,w,k,p", [(9, 8, 72, 10), (37, 53, 40, 20), (64, 130, 100, 100), (33, 257, 64, 64)])
def test_odd_shapes_exact_select_and_radius(cuda_device, dtype, h, w, k, p):
    _check_shape(cuda_device, dtype, h, w, k, p)


# TMA tile kernel: 16-byte-multiple row pitch (plain rows: 264, 512, 520) and 8-byte-multiple pitch
# (row pairs: 12, 20, 252, 260, 516, 620 -- last panels of 4, 4 and 108 valid columns, strips whose
# height has to be rounded to an even number of rows at 106 x 20)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("h,w,k,p", [(16, 12, 50, 50), (106, 20, 100, 64), (10, 252, 64, 64), (34, 260, 100, 100),
                                     (20, 516, 100, 100), (40, 620, 100, 100), (18, 264, 100, 100), (12, 512, 64, 64),
                                     (7, 520, 100, 100)])
def test_tile_kernel_shapes(cuda_device, dtype, h, w, k, p):
    want_path = "tile" if w % 8 == 0 else "tile_row_pairs"
    _check_shape(cuda_device, dtype, h, w, k, p, modes=("ties", "noise"), want_path=want_path)


@pytest.mark.parametrize("dtype", DTYPES)
def test_kernel_selection_for_reduced_precision(cuda_device, dtype):
    """Which peaks kernel runs (host-side query through the C ABI): cfg3/cfg5 maps (W = 612, 1224-byte
    rows) take TMA tiles over row pairs, 16-byte-multiple pitches plain tiles, the rest the per-lane kernel."""
    def path(h, w, **kw):
        cfg = DecodeConfig("sel", 1, 2, 1, h, w, 10, 10, cfg_id=3)
        raw = torch.zeros(1, 7, h, w, dtype=dtype, device=cuda_device)
        return ops.peaks_path(split_outputs(raw, cfg.labels, cfg.parts), 10, 10, **kw)
    assert path(512, 612) == "tile_row_pairs"
    assert path(128, 128) == "tile" and path(256, 256) == "tile"
    assert path(511, 612) == "warp"            # odd H: no row pairs
    assert path(512, 612, radius=1) == "warp"   # row pairs need tiles that start on an even row
    assert path(64, 130) == "warp" and path(37, 53) == "warp"
    assert path(512, 612, warp_kernel=True) == "warp"


def _check_shape(cuda_device, dtype, h, w, k, p, modes=("ties",), want_path=None):
    for mode in modes:
        cfg = DecodeConfig("oddhalf", 2, 3, 2, h, w, k, p, cfg_id=19)
        raw = make_raw(cfg, mode).to(dtype)
        outs = split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts)
        if want_path is not None:
            assert ops.peaks_path(outs, k, p) == want_path
        f32 = raw.float().numpy()
        for kw, okw in (({}, {}), ({"exact_select": True}, {}), ({"radius": 1}, {"radius": 1}), ({"warp_kernel": True}, {})):
            got = packed_np(ops.decode_packed(outs, k, p, 0.4, 0.1, **kw))
            want = O.decode_packed(f32[:, :3], f32[:, 3:5], f32[:, 5:7], f32[:, 7:], k, p, 0.4, 0.1,
                                   activation_fn=_device_activation(dtype, cuda_device),
                                   conf_cmp=float(torch.tensor(0.4, dtype=dtype)), **okw)
            assert_packed_equal(got, want, what=f"{dtype} {h}x{w} {mode} {kw}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_matches_reference_op_sequence_on_device(cuda_device, dtype):
    """The reference's own ops on reduced-precision CUDA tensors (k > 32: canonical topk order)."""
    cfg = CONFIGS["cfg3"]
    raw = make_raw(cfg, "noise", batch=3).to(cuda_device).to(dtype)
    outs = split_outputs(raw, cfg.labels, cfg.parts)
    got = ops.decode_packed(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
    ref = TP.decode_tensors(outs, cfg.max_objects, cfg.max_parts, cfg.conf_threshold, cfg.dist_thresh)
    for key in ("anchor_inds", "part_inds", "assign"):
        assert torch.equal(getattr(got, key), ref[key].to(getattr(got, key).dtype)), key
    assert torch.equal(got.anchor_out, ref["anchor_out"].float()) and torch.equal(got.part_out, ref["part_out"].float())


@pytest.mark.parametrize("dtype", DTYPES)
def test_drop_in_decoder_with_reduced_precision(cuda_device, dtype):
    cfg = CONFIGS["cfg1"]
    raw = make_raw(cfg, "blobs", batch=3).to(dtype)
    outs = split_outputs(raw.to(cuda_device), cfg.labels, cfg.parts)
    args = make_args(cfg)
    meta = Decoder(args)(outs, return_metadata=True)
    assert meta["anchor_hm_sig"].dtype == dtype and meta["embeddings"].dtype == dtype
    f32 = raw.float().numpy()
    want = O.decode_packed(f32[:, :2], f32[:, 2:3], f32[:, 3:5], f32[:, 5:], cfg.max_objects, cfg.max_parts,
                           cfg.conf_threshold, cfg.dist_thresh, activation_fn=_device_activation(dtype, cuda_device),
                           conf_cmp=float(torch.tensor(cfg.conf_threshold, dtype=dtype)))
    objs = O.assemble(want, args._r_labels, args._r_parts, args.anchor_name, cfg.conf_threshold,
                      (cfg.width, cfg.height), (4 * cfg.width, 4 * cfg.height))
    assert plain(meta["annotation"]) == objs
    np.testing.assert_array_equal(meta["topk_anchor"][0].cpu().numpy(), want["anchor_scores_masked"])


@pytest.mark.parametrize("dtype,w,wp,path", [
    (torch.float32, 128, 132, "tile"), (torch.float32, 300, 308, "tile"),
    (torch.float16, 256, 264, "tile"), (torch.bfloat16, 264, 272, "tile"),
    (torch.float16, 260, 268, "tile_row_pairs"), (torch.bfloat16, 252, 260, "tile_row_pairs"),
    (torch.float16, 516, 524, "tile_row_pairs"), (torch.float16, 612, 620, "tile_row_pairs"),
    (torch.float16, 130, 140, "warp"),
])
def test_padded_row_pitch_never_leaks(cuda_device, dtype, w, wp, path):
    """Row pitch larger than W (views of a wider allocation): the padding holds +30 logits that would win
    every top-K if a tile ever let them in -- the row-pair tiles physically read them and must blank them."""
    cfg = DecodeConfig("padded", 2, 2, 1, 34, w, 100, 100, cfg_id=23)
    raw = make_raw(cfg, "noise").to(dtype).to(cuda_device)
    wide = torch.full((raw.shape[0], raw.shape[1], cfg.height, wp), 30.0, dtype=dtype, device=cuda_device)
    wide[..., :w] = raw
    outs_pad = split_outputs(wide[..., :w], cfg.labels, cfg.parts)
    assert outs_pad["anchor_hm"].stride(2) == wp
    assert ops.peaks_path(outs_pad, 100, 100) == path
    got = ops.decode_packed(outs_pad, 100, 100, 0.4, 0.1)
    want = ops.decode_packed(split_outputs(raw, cfg.labels, cfg.parts), 100, 100, 0.4, 0.1, warp_kernel=True)
    torch.cuda.synchronize()
    n = got.blob.numel() - got.diag.numel() * 4
    assert torch.equal(got.blob[:n], want.blob[:n])
    assert float(got.anchor_out[..., 2].max()) < 0.9999  # no +30 logit (score 1 - 1e-6) got in
