"""Generate tests/golden/*.npz by executing the UNMODIFIED reference in this container.

    python -m oracle.make_golden            (needs /root/reference; writes small fixtures)

FPS indices are recovered from the reference's row output with the index-column trick
(only pc[:, :3] is used for distances, utils/utils.py:894). NN fixtures come from the
reference modules (pointNet/model/pointnetAtt.py) with torch.manual_seed-initialised weights.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def fps_cases():
    rng = np.random.default_rng(20261018)
    cases = {}
    cases["uniform_f32"] = (rng.random((1500, 5), dtype=np.float32), 256)
    cases["dups_f32"] = (np.repeat(rng.random((120, 4), dtype=np.float32), 8, axis=0), 300)
    g = np.stack(np.meshgrid(np.arange(9), np.arange(9), np.arange(9), indexing="ij"), -1).reshape(-1, 3)
    cases["grid_f32"] = (g.astype(np.float32), 200)
    cases["collinear_f32"] = (np.stack([np.linspace(0, 1, 700, dtype=np.float32)] * 3, 1), 128)
    cases["all_points_f32"] = (rng.random((97, 3), dtype=np.float32), 97)
    cases["uniform_f64"] = (rng.random((800, 6)), 128)
    cl = np.concatenate([rng.normal(c, 0.01, (300, 3)) for c in ((0, 0, 0), (1, 1, 0), (0, 1, .2))])
    cases["clustered_f32"] = (np.concatenate([cl, rng.random((900, 8))], 1).astype(np.float32), 384)
    return cases


def make_fps(uu):
    out = {}
    for name, (pc, S) in fps_cases().items():
        aug = np.concatenate([pc, np.arange(len(pc), dtype=pc.dtype)[:, None]], axis=1)
        rows = uu.fps(aug, S)
        idx = rows[:, -1].astype(np.int64)
        assert (aug[idx] == rows).all()
        out[name + "__pc"] = pc
        out[name + "__idx"] = idx
    np.savez_compressed(os.path.join(GOLD, "fps_reference.npz"), **out)
    print("fps_reference.npz:", len(out) // 2, "cases")


def main():
    os.makedirs(GOLD, exist_ok=True)
    model, uu, coll = ref_import.load()
    make_fps(uu)
    try:
        from oracle import make_golden_nn
        make_golden_nn.make(model, GOLD)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
