"""Multi-GPU plumbing: one process per GPU, videos sharded across ranks, NO collective on the data
path (per-video work is independent; reference src/batch_processor.cpp:215-235 deals files from one
queue). torch.distributed is used only for the barrier around the timed region, max-over-ranks of the
timings, sums of the work counters and gathering the small per-video results on rank 0."""
from __future__ import annotations

import os


class Dist:
    def __init__(self, backend: str | None = None, device=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._pg = False
        self._device = device
        if self.world > 1:
            import torch.distributed as dist

            self._dist = dist
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)
                self._pg = True
            self.backend = dist.get_backend()
        else:
            self.backend = None

    def _tensor(self, x):
        import torch

        dev = self._device if (self.backend == "nccl" and self._device is not None) else "cpu"
        return torch.tensor([float(x)], dtype=torch.float64, device=dev)

    def barrier(self):
        if self.world > 1:
            self._dist.barrier()

    def allmax(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self._tensor(x)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self._tensor(x)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return float(t.item())

    def gather_objects(self, obj):
        """list of every rank's obj on rank 0 (None elsewhere)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self._dist.gather_object(obj, out, dst=0)
        return out

    def close(self):
        if self._pg:
            self._dist.destroy_process_group()
            self._pg = False


def shard_videos(n_videos: int, world: int, rank: int) -> list[int]:
    """Round-robin deal of video indices — what a shared file queue converges to for equal-cost
    videos; disjoint across ranks and covering [0, n_videos)."""
    return list(range(rank, n_videos, world))


def throughput(units_per_rank_sum: float, steps: int, max_ms_over_ranks: float) -> float:
    """Whole-job units/s: everything all ranks processed ÷ the slowest rank's time."""
    return units_per_rank_sum * steps / (max_ms_over_ranks * 1e-3)


def strong_share(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Strong scaling of ONE stream: rank g scans frames [g·F/G, (g+1)·F/G) — contiguous,
    disjoint, covering, sizes within one frame of each other. Returns (first_frame, n)."""
    a = n_frames * rank // world
    return a, n_frames * (rank + 1) // world - a


def video_pieces(video_starts, first_frame: int, n: int):
    """Local frame offsets at which videos (or pieces of videos cut by a share's ends) begin inside the share
    [first_frame, first_frame + n), plus n: what K-C receives as video boundaries. video_starts: global start frames."""
    inner = sorted({int(s) - first_frame for s in video_starts if first_frame < int(s) < first_frame + n})
    return [0] + inner + [n]
