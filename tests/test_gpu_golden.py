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
    # host-side frame selection (motion_scanner.cpp:303-371) is input preparation here; the C++ host
    # mirror's own selection is checked in test_host_cli.py
    cnt1, _, recs1, pts1 = c.subset(c.selected(chunked=False))
    cnt2, _, recs2, pts2 = c.subset(c.selected(chunked=True))
    with ms.Context(0, c.params) as ctx:
        ctx.video_open(1, c.width, c.height)
        ctx.submit(1, pts1, cnt1, recs1 if len(recs1) else None)
        flags1, _ = ctx.collect(1)
        ctx.video_open(2, c.width, c.height)  # frames arrive in chunk order, as the pipeline's workers deliver them
        ctx.submit(2, pts2, cnt2, recs2 if len(recs2) else None)
        job, res = ctx.segments(2, e["duration"])
    assert pts1[flags1.astype(bool)].tobytes() == e["ts"].tobytes()
    assert res.decision == e["decision"]
    assert np.stack([job["start"], job["end"]], 1).reshape(-1, 2).tobytes() == e["segs"].tobytes()
    if res.decision != ms.NO_MOTION:
        assert np.float64(res.time_removed).tobytes() == np.float64(e["time_removed"]).tobytes()
        assert np.float64(res.saved_pct).tobytes() == np.float64(e["saved_pct"]).tobytes()
