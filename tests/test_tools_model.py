"""tools/lz4_tile_model.c (the sequential CPU twin of the GPU compressor's parse) must emit valid LZ4 and stay close to
the reference's compression ratio: every frame is decoded with the strict oracle decoder and with liblz4."""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
import oracle.reflib as ref
from skyplane_b200 import synth

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tools import tile_model as tm  # noqa: E402


@pytest.mark.parametrize("opts", [tm.kernel_opts(), tm.kernel_opts(3072, 512), tm.kernel_opts(2048, 256, 0), tm.Opts(4096, 1024, 4, 0, 0)])
def test_model_emits_valid_lz4(opts):
    rng = np.random.default_rng(3)
    datas = [rng.bytes(n) for n in (1, 12, 13, 100, 65536)] + [bytes(70000), (b"abcdefg" * 20000)[:131073], b"",
                                                                synth.silesia_like_chunk(11, 300000), b"x" * 13 + rng.bytes(40) + b"x" * 200,
                                                                rng.bytes(40000) + synth.silesia_like_chunk(3, 90000)]
    for d in datas:
        fr = tm.frame(d, opts)
        assert oracle.lz4f_decode(fr, len(d)) == d
        assert ref.lz4f_decompress(fr, len(d)) == d
        assert len(fr) <= 15 + len(d) + 4 * ((len(d) + 65535) // 65536) + 4


def test_kernel_parse_ratio_close_to_reference():
    """The shipped parse (4096 entries, 5-byte hash, backward extension) keeps >= 0.965 x the reference's ratio."""
    datas = [synth.silesia_like_chunk(i, 4 << 20) for i in range(2)]
    ours = sum(len(tm.frame(d)) for d in datas)
    refsz = sum(len(ref.lz4f_compress(d)) for d in datas)
    assert refsz / ours >= 0.965


def test_in_group_candidates_cover_short_periods():
    """The kernel's 32 slots of a group look the table up before any of them is inserted, so repeats closer than a group
    (32-bit words, doubles, pixels, runs) would be invisible to the table alone.  The in-group rule -- equal 32-bit hash 3, 4
    or 8 slots back -- must keep such data within 3 % of what fully sequential probing finds, and the frames stay valid."""
    rng = np.random.default_rng(9)
    walk = np.cumsum(rng.choice(np.array([0, 0, 0, 0, 1, -1, 2]), size=1 << 17)) + 1000
    kinds = {
        "int32 walk": walk.astype("<i4").tobytes(),
        "float64 walk": (walk * 0.25).astype("<f8").tobytes(),
        "rgb runs": np.repeat(rng.integers(0, 255, size=(1 << 10, 3), dtype=np.uint8), 64, axis=0).tobytes(),
        "utf-16 text": synth.silesia_like_chunk(5, 200000).decode("latin1").encode("utf-16-le"),
    }
    sequential = tm.Opts(4096, 1024, 4, 1, 1, 0, 0)   # every slot sees the slot before it
    table_only = tm.Opts(4096, 1024, 4, 1, 1, 1, 0)   # group-wise lookups, no in-group candidates
    sizes = {}
    for name, d in kinds.items():
        fr = tm.frame(d)
        assert oracle.lz4f_decode(fr, len(d)) == d and ref.lz4f_decompress(fr, len(d)) == d
        sizes[name] = (len(fr), len(tm.frame(d, sequential)), len(tm.frame(d, table_only)))
        assert sizes[name][0] <= 1.03 * sizes[name][1], (name, sizes[name])
    ours, _, none = sizes["int32 walk"]
    assert none >= 1.25 * ours, sizes["int32 walk"]  # what the rule is there for
