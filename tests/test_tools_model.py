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
