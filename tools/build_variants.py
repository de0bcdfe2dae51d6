#!/usr/bin/env python
"""Cross-compile tuning builds of libskychunk.so into tools/bin/ (they travel to the GPU box; SKYCHUNK_LIB selects one)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skyplane_b200 import build  # noqa: E402

VARIANTS = {
    "p10_e4096": {"SKY_PARSERS": 10, "SKY_LZ4_ENTRIES": 4096},
    "p8_e4096": {"SKY_PARSERS": 8, "SKY_LZ4_ENTRIES": 4096},
    "p6_e4096": {"SKY_PARSERS": 6, "SKY_LZ4_ENTRIES": 4096},
    "p12_e3072": {"SKY_PARSERS": 12, "SKY_LZ4_ENTRIES": 3072},
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    out = ROOT / "tools" / "bin"
    out.mkdir(exist_ok=True)
    for n in names:
        print(build.build_variant(out / f"libskychunk_{n}.so", VARIANTS[n]))
