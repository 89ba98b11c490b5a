"""Recipe for ``oracle/_ref/``: an importable archive of the UNMODIFIED reference package -- TEST INFRASTRUCTURE.

The reference (laclouis5/StructureDetector) is a pure-Python package whose build backend (hatchling) is not
installed here, so ``pip install --target`` cannot run; this script does what that install would do for a
pure-Python wheel: it packs ``/root/reference/src/sdnet`` -- byte for byte, nothing edited -- into
``oracle/_ref/sdnet_reference.zip`` (Python imports straight from a zip on ``sys.path``).  ``oracle/_ref/`` is
git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot, so
``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference's OWN ``Decoder`` there
(``kind: "reference"``); without the archive they fall back to ``oracle/torch_port.py`` (``kind: "port"``).

    python -m oracle.build_ref          # no-op (exit 0) when /root/reference is absent

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may use this module.
"""
from __future__ import annotations

import sys
import zipfile
from pathlib import Path

REFERENCE_SRC = Path("/root/reference/src")
REF_DIR = Path(__file__).resolve().parent / "_ref"
ARCHIVE = REF_DIR / "sdnet_reference.zip"


def build_ref(verbose: bool = True) -> Path | None:
    pkg = REFERENCE_SRC / "sdnet"
    if not pkg.is_dir():
        if verbose:
            print(f"{pkg} not present: keeping whatever is in {REF_DIR}")
        return ARCHIVE if ARCHIVE.exists() else None
    REF_DIR.mkdir(exist_ok=True)
    files = sorted(p for p in pkg.rglob("*.py"))
    with zipfile.ZipFile(ARCHIVE, "w", compression=zipfile.ZIP_DEFLATED) as zf:
        for path in files:
            info = zipfile.ZipInfo(str(path.relative_to(REFERENCE_SRC)), date_time=(2020, 1, 1, 0, 0, 0))
            info.compress_type = zipfile.ZIP_DEFLATED
            zf.writestr(info, path.read_bytes())
    if verbose:
        print(f"packed {len(files)} reference modules into {ARCHIVE}")
    return ARCHIVE


def load_reference_decoders():
    """The reference's ``sdnet.data.decoders`` module, imported from the archive (or, in the build container,
    from /root/reference itself); ``None`` when neither exists or its own imports fail."""
    sys.dont_write_bytecode = True
    source = ARCHIVE if ARCHIVE.exists() else (REFERENCE_SRC if (REFERENCE_SRC / "sdnet").is_dir() else None)
    if source is None:
        return None
    if str(source) not in sys.path:
        sys.path.insert(0, str(source))
    try:
        import sdnet.data.decoders as ref_decoders  # noqa: WPS433

        return ref_decoders
    except Exception as exc:  # noqa: BLE001 -- a missing third-party import on this box: report, fall back to the port
        print(f"[oracle] reference package not importable ({exc!r}); using the port", file=sys.stderr)
        return None


if __name__ == "__main__":
    build_ref()
