"""GPU: the CUDA path through the C ABI against outputs of the REFERENCE's own code
(tests/golden/ref_golden.json) — timestamps with motion, FFmpegJob segments, decision, savings."""
import numpy as np
import pytest

import motionscan as ms
from test_ref_golden import GOLDEN, cases, expected

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_gpu_matches_reference_outputs(name):
    c, e = cases()[name], expected(name)
    with ms.Context(0, c.params) as ctx:
        ctx.video_open(1, c.width, c.height)
        ctx.submit(1, c.pts, c.cnt, c.recs if len(c.recs) else None)
        flags, counts = ctx.collect(1)
        job, res = ctx.segments(1, e["duration"])
    assert c.pts[flags.astype(bool)].tobytes() == e["ts"].tobytes()
    assert res.decision == e["decision"]
    assert np.stack([job["start"], job["end"]], 1).reshape(-1, 2).tobytes() == e["segs"].tobytes()
    if res.decision != ms.NO_MOTION:
        assert np.float64(res.time_removed).tobytes() == np.float64(e["time_removed"]).tobytes()
        assert np.float64(res.saved_pct).tobytes() == np.float64(e["saved_pct"]).tobytes()
