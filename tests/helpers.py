"""Shared test helpers: seeded graphs with the edge cases the domain has (self loops, duplicate
edges, isolated nodes, hubs that exceed the long-row threshold, empty graphs)."""
import glob
import os

import numpy as np
import torch

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_golden(path):
    z = np.load(path)
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def random_graph(n, e, seed, self_loops=0, dups=0, hub=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(n, (e,), generator=g)
    dst = torch.randint(n, (e,), generator=g)
    if self_loops:
        l = torch.randint(n, (self_loops,), generator=g)
        pos = torch.randperm(src.numel() + self_loops, generator=g)
        src, dst = torch.cat([src, l])[pos], torch.cat([dst, l])[pos]
    if dups:
        src, dst = torch.cat([src, src[:dups]]), torch.cat([dst, dst[:dups]])
    if hub:                                   # node 1 receives `hub` edges, node 2 sends `hub` edges
        hs = torch.randint(n, (hub,), generator=g)
        src = torch.cat([src, hs, torch.full((hub,), 2)])
        dst = torch.cat([dst, torch.full((hub,), 1), hs])
    return torch.stack([src, dst]).long()


CASES = {
    "tiny": lambda: (random_graph(5, 7, 0), 5),
    "loops_dups": lambda: (random_graph(50, 400, 1, self_loops=9, dups=30), 50),
    "isolated": lambda: (random_graph(200, 150, 2), 260),          # ids 200..259 never appear
    "hub": lambda: (random_graph(300, 2000, 3, self_loops=4, hub=6000), 300),
    "empty": lambda: (torch.zeros((2, 0), dtype=torch.long), 7),
    "single_node": lambda: (torch.zeros((2, 3), dtype=torch.long), 1),
    "medium": lambda: (random_graph(5000, 60000, 4, self_loops=50, dups=500, hub=3000), 5000),
}


def relerr(a, b):
    """norm-wise error ||a-b||_inf / ||b||_inf."""
    a, b = a.double().cpu(), b.double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
