#!/usr/bin/env python
"""Cross-compile tuning builds of libskychunk.so into tools/bin/ (they travel to the GPU box; SKYCHUNK_LIB selects one)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skyplane_b200 import build  # noqa: E402

VARIANTS = {
    "e4096_w13": {"SKY_LZ4_ENTRIES": 4096, "SKY_WARPS": 13},
    "e4096_w12": {"SKY_LZ4_ENTRIES": 4096, "SKY_WARPS": 12},
    "e3072_w18": {"SKY_LZ4_ENTRIES": 3072, "SKY_WARPS": 18},
    "e2048_w26": {"SKY_LZ4_ENTRIES": 2048, "SKY_WARPS": 26},
    "e2048_w16": {"SKY_LZ4_ENTRIES": 2048, "SKY_WARPS": 16},
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    out = ROOT / "tools" / "bin"
    out.mkdir(exist_ok=True)
    for n in names:
        print(build.build_variant(out / f"libskychunk_{n}.so", VARIANTS[n]))
