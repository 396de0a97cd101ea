"""Pair batcher: shards frame pairs across the GPUs of one box (one process per GPU).

The reference scatters the batch with nn.DataParallel (train.py:493); here every rank owns a
contiguous slice of the pair batch and runs the co-attention on it with no data-path collective
(every (pair, modality) is independent: the batch dim of the bmm's at rgbd_segmentation_RAA.py:160-170).
For test.py-style inference (test.py:278-305) the unit that is sharded is the QUERY frame, so that
the `sample_range` reference frames of one query and the mean over them (test.py:301-305) stay local.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple

import torch


def shard_range(num_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `num_units` over `world_size` ranks, remainder to the low ranks.

    Returns (start, count) of this rank's slice.
    """
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if num_units < 0:
        raise ValueError("num_units must be >= 0")
    base, rem = divmod(num_units, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


@dataclass
class PairBatch:
    """Encoder features of a batch of frame pairs for both modalities (fp32, [n, 256, H', W'])."""
    v_a: torch.Tensor
    v_b: torch.Tensor
    d_a: torch.Tensor
    d_b: torch.Tensor
    pair_ids: List[int]


class SyntheticPairBatcher:
    """Deterministic synthetic feature pairs, V = prelu_0.25(N(0,1)) * sigma.

    Pair `i` of the global batch is generated from seed (base_seed, i) regardless of which rank owns
    it, so that a sharded run produces exactly the tensors of the single-process run.
    """

    def __init__(self, num_pairs: int, feat_hw: Tuple[int, int], sigma: float = 0.66, channels: int = 256,
                 base_seed: int = 1234, world_size: int = 1, rank: int = 0, device: str | torch.device = "cpu",
                 pin_memory: bool = False):
        self.num_pairs = num_pairs
        self.h, self.w = feat_hw
        self.sigma = sigma
        self.c = channels
        self.base_seed = base_seed
        self.world_size, self.rank = world_size, rank
        self.start, self.count = shard_range(num_pairs, world_size, rank)
        self.device = torch.device(device)
        self.pin_memory = pin_memory

    def _one(self, pair_id: int) -> torch.Tensor:
        g = torch.Generator(device="cpu")
        g.manual_seed(self.base_seed * 1_000_003 + pair_id)
        x = torch.randn(4, self.c, self.h, self.w, generator=g, dtype=torch.float32)
        return torch.where(x >= 0, x, 0.25 * x) * self.sigma

    def local_pair_ids(self) -> List[int]:
        return list(range(self.start, self.start + self.count))

    def batches(self, batch_size: int) -> Iterator[PairBatch]:
        ids = self.local_pair_ids()
        for i in range(0, len(ids), batch_size):
            chunk = ids[i:i + batch_size]
            feats = torch.stack([self._one(p) for p in chunk], dim=1)  # [4, n, C, H, W]
            if self.pin_memory and self.device.type == "cpu":
                feats = feats.pin_memory()
            feats = feats.to(self.device, non_blocking=True)
            yield PairBatch(feats[0], feats[1], feats[2], feats[3], chunk)


def query_reference_groups(num_queries: int, sample_range: int, frames_per_sequence: int, seed: int = 1234):
    """test.py-style pairing (dataloaders/sbm_rgbd_loader.py:556-574): every query frame is paired with
    `sample_range` frames drawn without replacement from its own sequence (the query itself may be drawn).

    Returns a list of (query_frame, [reference_frames]) with frame indices local to the sequence.
    """
    import random
    rng = random.Random(seed)
    groups = []
    for q in range(num_queries):
        k = min(sample_range, frames_per_sequence)
        groups.append((q % frames_per_sequence, rng.sample(range(frames_per_sequence), k)))
    return groups


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    afterwards are node-local (first touch) and its H2D / D2H traffic does not cross the inter-socket link.  With one
    process per GPU (the launch model here) the unbound default puts every rank's buffers wherever the kernel scheduled
    it -- usually node 0 -- and the ranks then share one socket's memory controllers.  Best effort: returns what it did."""
    import os
    import torch
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(os.path.join(base, "numa_node")) as f:
            node = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info.update(pci=bdf, numa_node=node, cpus=len(cpus))
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:  # sysfs layout / permissions differ between hosts: never fatal
        info["error"] = f"{type(e).__name__}: {e}"
    return info
