"""-m gpu: dense sigmoid + NMS maps (the reference's RawDecoder, src/sdnet/cli/convert_coreml.py:12-19)
against the reference's own op sequence on the device, and the RawDecoder -> CoreMLDecoder route against
the plain Decoder."""
import pytest
import torch

from oracle import torch_port as TP
from structuredetector_b200 import CoreMLDecoder, Decoder, RawDecoder, ops
from structuredetector_b200.synth import CONFIGS, DecodeConfig, make_raw, split_outputs
from tests.helpers import make_args, plain

pytestmark = pytest.mark.gpu


def _reference_maps(hm, radius=2):
    return TP.suppress(TP.activate(hm), radius)  # max_pool2d equality on clamp(sigmoid(x)), in hm's dtype


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("h,w,mode", [(128, 128, "noise"), (128, 128, "ties"), (37, 53, "ties"), (1, 300, "noise"),
                                      (300, 1, "noise"), (33, 257, "blobs"), (64, 612, "noise")])
def test_suppress_maps_bit_exact(cuda_device, dtype, h, w, mode):
    cfg = DecodeConfig("sup", 2, 3, 2, h, w, 1, 1, cfg_id=31)
    raw = make_raw(cfg, mode).to(dtype).to(cuda_device)
    hm = raw[:, :5]  # channel-slice view, not contiguous in the batch dimension
    for radius in (2, 1):
        got = ops.suppress_maps(hm, radius)
        want = _reference_maps(hm, radius)
        assert got.dtype == dtype and got.shape == want.shape
        assert torch.equal(got.view(torch.int32 if dtype == torch.float32 else torch.int16),
                           want.contiguous().view(torch.int32 if dtype == torch.float32 else torch.int16)), (h, w, mode, radius)


def test_saturated_plateaus_survive(cuda_device):
    hm = torch.full((1, 1, 20, 24), 20.0, device=cuda_device)  # every score is the clamp: one big plateau
    hm[0, 0, 5, 5] = 15.0                                       # a lower logit with the SAME clamped score
    got = ops.suppress_maps(hm)
    assert torch.equal(got, _reference_maps(hm)) and int((got > 0).sum()) == 20 * 24


@pytest.mark.parametrize("name,mode", [("cfg1", "blobs"), ("cfg1", "ladder"), ("cfg3", "blobs")])
def test_raw_decoder_feeds_coreml_decoder(cuda_device, name, mode):
    cfg = CONFIGS[name]
    raw = make_raw(cfg, mode, batch=2).to(cuda_device)
    args = make_args(cfg)
    baked = RawDecoder(cfg.labels + cfg.parts)(raw)
    assert baked.shape == raw.shape and torch.equal(baked[:, cfg.labels + cfg.parts:], raw[:, cfg.labels + cfg.parts:])
    via_coreml = CoreMLDecoder(args)(split_outputs(baked, cfg.labels, cfg.parts))
    direct = Decoder(args)(split_outputs(raw, cfg.labels, cfg.parts))
    assert plain(via_coreml) == plain(direct)
