"""Deterministic synthetic network outputs for the SDNet decoding path.

The decoder consumes the four channel-slice views that the reference network's
forward returns (reference: src/sdnet/model/network.py:79-84): one NCHW tensor
``raw (B, M+N+4, H, W)`` sliced into anchor heat maps ``[:, :M]``, part heat maps
``[:, M:M+N]``, sub-pixel offsets ``[:, M+N:M+N+2]`` and part->anchor embeddings
``[:, M+N+2:]``.  Everything here is generated on the CPU from a seeded
``torch.Generator`` so that the CPU oracle and the GPU path see identical bits.

Modes
-----
``noise``   dense: heat logits 2*N(0,1)-3 (every channel saturates K peaks).
``blobs``   realistic: low background, Gaussian bumps for anchors and their parts,
            embedding planes pointing from each part to its anchor.
``ladder``  tie-free: like ``blobs`` but every above-background peak gets its own
            probability from a strictly decreasing ladder (gap >= 1e-4), peaks are
            kept >= 3 px apart, so no sigmoid implementation can reorder them.
``ties``    adversarial: saturated plateaus, duplicated logits, border/corner
            peaks, equidistant anchors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

MODES = ("noise", "blobs", "ladder", "ties")


@dataclass(frozen=True)
class DecodeConfig:
    """One row of BASELINE.json ``configs`` ("Concrete sizes")."""

    name: str
    batch: int
    labels: int  # M anchor classes
    parts: int  # N part kinds
    height: int  # map rows (input / 4)
    width: int  # map cols (input / 4)
    max_objects: int  # K
    max_parts: int  # P
    conf_threshold: float = 0.4
    dist_thresh: float = 0.1
    cfg_id: int = 0

    @property
    def channels(self) -> int:
        return self.labels + self.parts + 4

    @property
    def contract_bytes_per_image(self) -> int:
        """north_star's denominator: heat+offset+embedding planes read once (fp32)."""
        return self.channels * self.height * self.width * 4

    @property
    def min_bytes_per_image(self) -> int:
        """Heat maps once + one 32 B sector per gathered word + packed outputs."""
        k, p = self.max_objects, self.max_parts
        heat = (self.labels + self.parts) * self.height * self.width * 4
        return heat + 32 * 2 * (k + 2 * p) + 4 * (4 * k + 6 * p) + 8 * (k + p) + 4 * p


CONFIGS = {
    "cfg1": DecodeConfig("cfg1", 1, 2, 1, 128, 128, 100, 100, cfg_id=1),
    "cfg2": DecodeConfig("cfg2", 64, 2, 1, 128, 128, 100, 100, cfg_id=2),
    "cfg3": DecodeConfig("cfg3", 16, 2, 1, 512, 612, 100, 100, cfg_id=3),
    "cfg4": DecodeConfig("cfg4", 16, 20, 10, 256, 256, 500, 500, cfg_id=4),
    "cfg5": DecodeConfig("cfg5", 1024, 2, 1, 512, 612, 100, 100, cfg_id=5),
}


def split_outputs(raw: torch.Tensor, labels: int, parts: int) -> dict:
    """Channel-slice ``raw`` exactly like the producer network does (views, no copy)."""
    m, n = labels, parts
    return {
        "anchor_hm": raw[:, :m],
        "part_hm": raw[:, m : m + n],
        "offsets": raw[:, m + n : m + n + 2],
        "embeddings": raw[:, m + n + 2 : m + n + 4],
    }


def _logit(p: torch.Tensor) -> torch.Tensor:
    return torch.log(p / (1.0 - p))


def _stamp_bump(plane: torch.Tensor, cy: int, cx: int, peak_logit: float, sigma: float, floor: float):
    """max-combine a Gaussian-shaped logit bump centred on (cy, cx) into ``plane``."""
    h, w = plane.shape
    r = max(1, int(math.ceil(3 * sigma)))
    y0, y1 = max(0, cy - r), min(h, cy + r + 1)
    x0, x1 = max(0, cx - r), min(w, cx + r + 1)
    ys = torch.arange(y0, y1, dtype=torch.float32).unsqueeze(1) - cy
    xs = torch.arange(x0, x1, dtype=torch.float32).unsqueeze(0) - cx
    g = torch.exp(-(ys * ys + xs * xs) / (2.0 * sigma * sigma))
    bump = floor + (peak_logit - floor) * g
    plane[y0:y1, x0:x1] = torch.maximum(plane[y0:y1, x0:x1], bump)
    plane[cy, cx] = max(float(plane[cy, cx]), peak_logit)


def _structured(cfg: DecodeConfig, batch: int, gen: torch.Generator, ladder: bool) -> torch.Tensor:
    m, n, h, w = cfg.labels, cfg.parts, cfg.height, cfg.width
    raw = torch.empty(batch, cfg.channels, h, w, dtype=torch.float32)
    if ladder:
        # strictly sub-threshold background made of pairwise distinct, well separated logits
        # (a random permutation of an arithmetic ladder) so that not even the below-threshold
        # top-k slots can tie
        for b in range(batch):
            for c in range(m + n):
                perm = torch.randperm(h * w, generator=gen).double()
                steps = float(h * w * (m + n))  # channel-interleaved rungs: distinct across channels too
                raw[b, c] = (-10.0 + 4.0 * (perm * (m + n) + c) / steps).float().view(h, w)
    else:
        raw[:, : m + n] = -6.0 + 0.5 * torch.randn(batch, m + n, h, w, generator=gen)
    raw[:, m + n : m + n + 2] = torch.rand(batch, 2, h, w, generator=gen)
    raw[:, m + n + 2 :] = 0.3 * torch.randn(batch, 2, h, w, generator=gen)
    sigma = max(0.8, 0.1 * min(h, w) / 3.0 / 4.0) if not ladder else 0.7
    reach = max(4.0, 0.08 * min(h, w))
    step = 1e-4
    for b in range(batch):
        taken = torch.zeros(h, w, dtype=torch.bool)
        n_obj = int(torch.randint(5, 21, (1,), generator=gen))
        rung = 0

        def place(cy, cx):
            if not (0 <= cy < h and 0 <= cx < w):
                return False
            y0, y1, x0, x1 = max(0, cy - 3), min(h, cy + 4), max(0, cx - 3), min(w, cx + 4)
            if ladder and bool(taken[y0:y1, x0:x1].any()):
                return False
            taken[cy, cx] = True
            return True

        for _ in range(n_obj):
            cy = int(torch.randint(0, h, (1,), generator=gen))
            cx = int(torch.randint(0, w, (1,), generator=gen))
            if not place(cy, cx):
                continue
            cls = int(torch.randint(0, m, (1,), generator=gen))
            if ladder:
                p = 0.97 - step * 37 * rung
                rung += 1
            else:
                p = float(0.45 + 0.5 * torch.rand(1, generator=gen))
            _stamp_bump(raw[b, cls], cy, cx, float(_logit(torch.tensor(p))), sigma, -9.0 if ladder else -6.0)
            n_parts = int(torch.randint(1, 7, (1,), generator=gen))
            for _ in range(n_parts):
                dy = int((torch.rand(1, generator=gen) * 2 - 1) * reach)
                dx = int((torch.rand(1, generator=gen) * 2 - 1) * reach)
                py, px = cy + dy, cx + dx
                if not place(py, px):
                    continue
                kind = int(torch.randint(0, n, (1,), generator=gen))
                if ladder:
                    p = 0.97 - step * 37 * rung
                    rung += 1
                else:
                    p = float(0.3 + 0.65 * torch.rand(1, generator=gen))
                _stamp_bump(raw[b, m + kind], py, px, float(_logit(torch.tensor(p))), sigma, -9.0 if ladder else -6.0)
                noise = 0.3 * torch.randn(2, generator=gen) if not ladder else torch.zeros(2)
                # embedding = anchor - part (reference: src/sdnet/data/transforms.py:181)
                raw[b, m + n + 2, py, px] = float(cx - px) + float(noise[0])
                raw[b, m + n + 3, py, px] = float(cy - py) + float(noise[1])
    return raw


def _ties(cfg: DecodeConfig, batch: int, gen: torch.Generator) -> torch.Tensor:
    """Inputs whose ties come from *identical logits*, so every sigmoid agrees on them."""
    m, n, h, w = cfg.labels, cfg.parts, cfg.height, cfg.width
    raw = torch.empty(batch, cfg.channels, h, w, dtype=torch.float32)
    # quantised background: lots of exact repeats at sub-threshold level
    raw[:, : m + n] = torch.randint(-40, -20, (batch, m + n, h, w), generator=gen).float() / 4.0
    raw[:, m + n : m + n + 2] = torch.randint(0, 4, (batch, 2, h, w), generator=gen).float() / 4.0
    raw[:, m + n + 2 :] = torch.randint(-8, 9, (batch, 2, h, w), generator=gen).float() / 2.0
    levels = torch.tensor([0.0, 1.0, 1.0, 2.5, 2.5, 2.5, 6.0, 14.0, 15.0, 20.0])
    for b in range(batch):
        for c in range(m + n):
            plane = raw[b, c]
            # corners + borders carry peaks (max_pool2d pads with -inf)
            for (y, x) in ((0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (0, w // 2), (h // 2, 0)):
                plane[y, x] = float(levels[int(torch.randint(0, len(levels), (1,), generator=gen))])
            # duplicated peak values scattered at random
            for _ in range(48):
                y = int(torch.randint(0, h, (1,), generator=gen))
                x = int(torch.randint(0, w, (1,), generator=gen))
                plane[y, x] = float(levels[int(torch.randint(0, len(levels), (1,), generator=gen))])
            # a saturated plateau (all members clamp to the same score and all survive NMS)
            y = int(torch.randint(0, max(1, h - 7), (1,), generator=gen))
            x = int(torch.randint(0, max(1, w - 9), (1,), generator=gen))
            plane[y : y + 6, x : x + 8] = 16.0
            # a mid-value plateau
            y = int(torch.randint(0, max(1, h - 5), (1,), generator=gen))
            x = int(torch.randint(0, max(1, w - 5), (1,), generator=gen))
            plane[y : y + 4, x : x + 4] = 3.0
        # equidistant anchors: zero offsets/embeddings around a symmetric pattern
        if h >= 24 and w >= 24:
            cy, cx = h // 2, w // 2
            raw[b, m + n :, cy - 10 : cy + 11, cx - 10 : cx + 11] = 0.0
            raw[b, : m + n, cy - 10 : cy + 11, cx - 10 : cx + 11] = -10.0
            raw[b, 0, cy, cx - 6] = 5.0
            raw[b, 0, cy, cx + 6] = 5.0
            raw[b, m, cy, cx] = 5.0  # a part exactly between two equal anchors
    return raw


def make_raw(cfg: DecodeConfig, mode: str = "noise", batch: int | None = None, seed: int | None = None) -> torch.Tensor:
    """Return the contiguous ``raw (B, M+N+4, H, W)`` fp32 CPU tensor for ``cfg``/``mode``."""
    if mode not in MODES:
        raise ValueError(f"unknown synthetic mode {mode!r}; expected one of {MODES}")
    batch = cfg.batch if batch is None else batch
    gen = torch.Generator().manual_seed((1234 + cfg.cfg_id) if seed is None else seed)
    m, n, h, w = cfg.labels, cfg.parts, cfg.height, cfg.width
    if mode == "noise":
        raw = torch.empty(batch, cfg.channels, h, w, dtype=torch.float32)
        raw[:, : m + n] = 2.0 * torch.randn(batch, m + n, h, w, generator=gen) - 3.0
        raw[:, m + n : m + n + 2] = torch.rand(batch, 2, h, w, generator=gen)
        raw[:, m + n + 2 :] = 5.0 * torch.randn(batch, 2, h, w, generator=gen)
        return raw
    if mode == "ties":
        return _ties(cfg, batch, gen)
    return _structured(cfg, batch, gen, ladder=(mode == "ladder"))
