#!/usr/bin/env python
"""Cross-compile tuning builds of libskychunk.so into tools/bin/ (they travel to the GPU box; SKYCHUNK_LIB selects one)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skyplane_b200 import build  # noqa: E402

VARIANTS = {
    "pr2_pa10": {"SKY_PROBERS": 2, "SKY_PARSERS": 10},
    "pr3_pa9": {"SKY_PROBERS": 3, "SKY_PARSERS": 9},
    "pr4_pa8": {"SKY_PROBERS": 4, "SKY_PARSERS": 8},
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    out = ROOT / "tools" / "bin"
    out.mkdir(exist_ok=True)
    for n in names:
        print(build.build_variant(out / f"libskychunk_{n}.so", VARIANTS[n]))
