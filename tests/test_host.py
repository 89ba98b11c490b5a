"""CPU: everything on the host side of the boundary -- the C-ABI library loads and exports every
symbol include/sdnet_decode.h declares, argument errors are reported before any launch, the
product path refuses to run without a CUDA device, and the annotation types behave like the
reference's."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from structuredetector_b200 import ImageAnnotation, Keypoint, Object, _native, ops
from structuredetector_b200.annotations import Box
from structuredetector_b200.synth import CONFIGS, make_raw, split_outputs

ROOT = Path(__file__).resolve().parent.parent


def _header_functions():
    text = (ROOT / "include" / "sdnet_decode.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdnet_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/sdnet_decode.h but not exported"
    assert sorted(_native.EXPORTS) == declared
    assert lib.sdnet_abi_version() == _native.ABI_VERSION


def test_struct_layout_matches_the_header():
    # the library checks struct_size itself; a mismatch comes back as SDNET_E_STRUCT (-7)
    p = _native.SdnetDecodeParams()
    p.struct_size = ctypes.sizeof(_native.SdnetDecodeParams) - 8
    assert _native.load().sdnet_decode_launch(ctypes.byref(p), None) == -7
    assert "struct_size" in _native.error_string(-7)


def test_every_field_offset_matches_the_header_under_gcc(tmp_path):
    """The ctypes mirror of every struct of the C ABI, field by field: a C program that includes the public header
    prints sizeof and offsetof as the C compiler sees them."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {n: getattr(_native, n) for n in
               ("SdnetTensor4", "SdnetDecodeParams", "SdnetSchedule", "SdnetMatchParams", "SdnetObjectMatchParams")}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "sdnet_decode.h"', "int main(void) {"]
    for name, struct in structs.items():
        src.append(f'printf("{name} %zu\\n", sizeof({name}));')
        src += [f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));' for field, _ in struct._fields_]
    src.append("return 0; }")
    (tmp_path / "layout.c").write_text("\n".join(src))
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(tmp_path / "layout.c"), "-o", str(tmp_path / "layout")],
                   check=True)
    out = subprocess.run([str(tmp_path / "layout")], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in filter(None, out):
        name, value = line.split()
        if "." in name:
            struct, field = name.split(".")
            assert getattr(structs[struct], field).offset == int(value), name
        else:
            assert ctypes.sizeof(structs[name]) == int(value), name
        seen += 1
    assert seen == sum(len(s._fields_) + 1 for s in structs.values())


def test_argument_errors_are_reported_before_any_launch():
    lib = _native.load()
    p = _native.SdnetDecodeParams()
    p.struct_size = ctypes.sizeof(_native.SdnetDecodeParams)
    p.dtype, p.radius = 0, 2
    p.B, p.M, p.N, p.H, p.W, p.K, p.P = 1, 2, 1, 8, 8, 65, 10  # K > H*W: torch.topk's "k out of range"
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -2
    p.K = 10
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -1  # NULL tensors
    p.dtype = 3
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -4
    p.dtype, p.radius = 0, 3
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -6
    assert lib.sdnet_decode_launch(None, None) == -1
    # fused-gather destinations: unknown mode, multicast with more than its one (multicast) destination, too many peers
    p.radius = 2
    p.dest_mode, p.n_dest = 2, 1
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -2
    p.dest_mode, p.n_dest = _native.DEST_MULTICAST, 2
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -2
    p.dest_mode, p.n_dest = _native.DEST_PEER_STORES, _native.MAX_DEST + 1
    assert lib.sdnet_decode_launch(ctypes.byref(p), None) == -2
    assert lib.sdnet_gather_wait_launch(None, 2, 1, None) == -1
    flag_words = (ctypes.c_uint32 * 16)()
    assert lib.sdnet_gather_wait_launch(flag_words, 0, 1, None) == -2
    assert lib.sdnet_gather_wait_launch(flag_words, _native.MAX_DEST + 1, 1, None) == -2
    out = ctypes.c_size_t(0)
    assert lib.sdnet_decode_workspace_bytes(1, 2, 1, 4096, 4096, 10, 10, 0, ctypes.byref(out)) == -2  # H*W >= 2^24
    # per-plane candidate lists hold 8 max(K, P) + 8192 records, but never less than the exact select's bitmap needs
    ws = _native.workspace_bytes(16, 2, 1, 512, 612, 100, 100)
    assert 16 * 3 * (100 + 2 + 512 * 612 // 64) * 8 < ws < 16 * 3 * 16384 * 8
    with pytest.raises(RuntimeError):
        _native.check(-2, "x")
    with pytest.raises(ValueError):
        _native.check(-5, "x")


def test_product_path_has_no_cpu_fallback():
    cfg = CONFIGS["cfg1"]
    outs = split_outputs(make_raw(cfg, "blobs", batch=1), cfg.labels, cfg.parts)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.decode_packed(outs, 100, 100, 0.4, 0.1)
    with pytest.raises(TypeError):
        ops.decode_packed({k: v.double() for k, v in outs.items()}, 100, 100, 0.4, 0.1)
    with pytest.raises(RuntimeError):
        ops.activate_maps(outs["anchor_hm"])


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "structuredetector_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f"{path} reaches into oracle/"
    for path in (ROOT / "structuredetector_b200" / "csrc").glob("*"):
        if path.suffix in (".cu", ".h", ".cuh"):
            assert "oracle" not in path.read_text().lower()


def test_packed_blob_layout_is_aligned_and_disjoint():
    B, K, P, C = 3, 7, 5, 4
    blob = torch.zeros(ops.packed_nbytes(B, K, P, C), dtype=torch.uint8)
    v = ops._carve(blob, B, K, P, C)
    assert v.anchor_out.shape == (B, K, 4) and v.part_out.shape == (B, P, 6) and v.assign.shape == (B, P)
    spans = []
    for t in (v.anchor_inds, v.part_inds, v.anchor_out, v.part_out, v.part_emb, v.assign, v.counts, v.diag):
        off = t.data_ptr() - blob.data_ptr()
        assert off % 16 == 0
        spans.append((off, off + t.numel() * t.element_size()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= blob.numel()


def test_annotation_types_follow_the_reference_api(tmp_path):
    kp = Keypoint("leaf", 10.0, 20.0, 0.9)
    assert kp.resized((100, 50), (200, 200)).x == 20.0 and kp.x == 10.0
    kp.resize((100, 50), (200, 200))
    assert (kp.x, kp.y) == (20.0, 80.0)
    obj = Object("bean", Keypoint("stem", 1.0, 2.0, 0.8), [kp], Box(3, 4, 1, 2))
    assert obj.nb_parts == 1 and obj.x == 1.0
    obj.x = 5.0
    assert obj.anchor.x == 5.0
    assert obj.box.standardized().x_min == 1 and obj.box.width == 2 and obj.box.x_mid == 2
    rep = obj.json_repr()
    assert rep["label"] == "bean" and rep["parts"][0]["kind"] == "stem" and rep["parts"][1]["location"] == {"x": 20.0, "y": 80.0}
    back = Object.from_json(rep, "stem")
    assert back.anchor.score == 0.8 and back.parts[0].kind == "leaf" and back.box.json_repr() == obj.box.json_repr()
    ann = ImageAnnotation("batch_0", [obj], img_size=(100, 100))
    assert len(ann) == 1 and ann.nb_parts == 1 and not ann.is_empty and ann.image_name == "batch_0"
    norm = ann.normalized()
    assert norm.objects[0].x == 0.05 and ann.objects[0].x == 5.0
    ann.save_json(tmp_path)
    again = ImageAnnotation.from_json(tmp_path / "batch_0.json", "stem")
    assert again.objects[0].parts[0].score == 0.9 and again.img_size == [100, 100]
    assert json.loads((tmp_path / "batch_0.json").read_text())["objects"][0]["label"] == "bean"
    assert "Keypoint(kind: leaf" in repr(kp) and ImageAnnotation("x").is_empty


def _random_packed(rng, B, K, P, M, N, conf):
    """Packed rows the way the tail kernel leaves them: scores sorted descending, some exactly at float32(conf)
    (emitted by the double compare only when float32(conf) > conf), parts pointing at any slot or -1."""
    blob = torch.zeros(ops.packed_nbytes(B, K, P, M + N), dtype=torch.uint8)
    host = ops._carve(blob, B, K, P, M + N)
    a, p = host.anchor_out.numpy(), host.part_out.numpy()
    a[..., 0], a[..., 1] = rng.uniform(0, 600, (B, K)), rng.uniform(0, 500, (B, K))
    a[..., 2] = -np.sort(-rng.uniform(0.2, 1.0, (B, K)).astype(np.float32), axis=1)
    a[:, K // 2, 2] = np.float32(conf)
    a[..., 3] = rng.integers(0, M, (B, K))
    p[..., 0], p[..., 1] = rng.uniform(0, 600, (B, P)), rng.uniform(0, 500, (B, P))
    p[..., 2] = -np.sort(-rng.uniform(0.2, 1.0, (B, P)).astype(np.float32), axis=1)
    p[:, P // 3, 2] = np.float32(conf)
    p[..., 3] = rng.integers(0, N, (B, P))
    host.assign.numpy()[:] = rng.integers(-1, K, (B, P))
    return host


@pytest.mark.parametrize("conf", [0.4, 0.5, 0.25])
def test_c_object_assembly_matches_the_oracle(conf):
    """csrc/fastobj.c (what Decoder._assemble / _raw_parts run) against the numpy restatement of
    decoders.py:103-159: same objects, same order, same doubles -- including parts grouped onto an anchor
    that the double compare does not emit and the >= / > asymmetry between raw parts and objects."""
    from types import SimpleNamespace

    from oracle import sdnet_oracle as O
    from structuredetector_b200 import Decoder

    rng = np.random.default_rng(7)
    B, K, P, M, N = 5, 23, 31, 3, 2
    host = _random_packed(rng, B, K, P, M, N, conf)
    labels, kinds = {i: f"label{i}" for i in range(M)}, {i: f"part{i}" for i in range(N)}
    dec = Decoder(SimpleNamespace(_r_labels=labels, _r_parts=kinds, anchor_name="stem", down_ratio=4.0, max_objects=K,
                                  max_parts=P, conf_threshold=conf, decoder_dist_thresh=0.1))
    out_size, in_size = (612, 512), (2448, 2051)  # unequal, inexact ratios
    packed = {"anchor_out": host.anchor_out.numpy(), "part_out": host.part_out.numpy(), "assign": host.assign.numpy()}
    anns = dec._assemble(host, conf, out_size, in_size)
    want = O.assemble(packed, labels, kinds, "stem", conf, out_size, in_size)
    got = [[(o.name, (o.anchor.kind, o.anchor.x, o.anchor.y, o.anchor.score), [(k.kind, k.x, k.y, k.score) for k in o.parts])
            for o in ann.objects] for ann in anns]
    assert got == want
    assert [str(a.image_path) for a in anns] == [f"batch_{b}" for b in range(B)]
    assert all(o.box is None and isinstance(o.parts, list) for a in anns for o in a.objects)
    raw = dec._raw_parts(host, conf, out_size, in_size)
    assert [[(k.kind, k.x, k.y, k.score) for k in img] for img in raw] == O.raw_parts(packed, kinds, conf, out_size, in_size)
    # the instances behave like the reference's: mutable, deep-copyable, resizable
    first = anns[0].objects[0]
    x0 = first.anchor.x
    clone = anns[0].resized((2, 2), (1, 1))
    assert clone.objects[0].anchor.x == x0 / 2 and first.anchor.x == x0
    first.parts.append(Keypoint("part0", 1.0, 2.0, 0.5))
    assert first.nb_parts >= 1


@pytest.mark.parametrize("with_sizes", [False, True])
def test_packed_json_writer_equals_json_repr(tmp_path, with_sizes):
    """Decoder._json_from_host (csrc/fastobj.c json_text): the documents `detect` saves (cli/detect.py:41-52 ->
    utils.py:275-286), straight from the packed rows, byte for byte what json.dumps(annotation.json_repr(), indent=2)
    gives for the assembled (and, with image sizes, resized) annotations -- empty images, non-ASCII names included."""
    from types import SimpleNamespace

    from structuredetector_b200 import Decoder

    rng = np.random.default_rng(21)
    B, K, P, M, N = 4, 17, 23, 2, 2
    conf = 0.4
    host = _random_packed(rng, B, K, P, M, N, conf)
    host.anchor_out.numpy()[1, :, 2] = 0.1  # an image without objects: "objects": []
    host.anchor_out.numpy()[2, 0, :2] = (1e-7, 123456789.0)  # float repr corner cases
    labels, kinds = {0: "bean", 1: 'ma"ïze'}, {0: "leaf", 1: "feuille\\n"}
    dec = Decoder(SimpleNamespace(_r_labels=labels, _r_parts=kinds, anchor_name="stem", down_ratio=4.0, max_objects=K,
                                  max_parts=P, conf_threshold=conf, decoder_dist_thresh=0.1))
    out_size, in_size = (612, 512), (2448, 2048)
    paths = [tmp_path / f"img_{b}.jpg" for b in range(B)] if with_sizes else None
    sizes = [(4000 + 13 * b, 3000 + 7 * b) for b in range(B)] if with_sizes else None
    texts = dec._json_from_host(host, conf, out_size, in_size, paths, sizes, in_size if with_sizes else None)
    anns = dec._assemble(host, conf, out_size, in_size)
    for b, (ann, text) in enumerate(zip(anns, texts)):
        if with_sizes:  # what detect.py does with the decoder's annotation
            ann.resize(in_size, sizes[b])
            ann.img_size = sizes[b]
            ann.image_path = paths[b]
        assert text == json.dumps(ann.json_repr(), indent=2), f"image {b}"
        assert json.loads(text)["objects"] == ann.json_repr()["objects"]
