"""Generate tests/golden/heist_golden_r3.npz by running the UNMODIFIED Python reference (third fixture file).

Run in the build container only (the reference tree is not present on the GPU box):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_r3.py

* patrol*  - guards whose patrols exercise every case of Guard.update (security.py:145-159) and of reset()
             (environment.py:205-208: guards back to waypoint 0, headings kept): strides 2, 3, -1 and a multiple of
             the path length, paths with repeated waypoints (moves that are no move keep the heading), paths of
             one and two waypoints, up to four guards whose cones overlap, a camera next to them -- with short
             episodes (max_steps 7 ... 13, not multiples of the path lengths) so that resets fall on every phase of a
             patrol.  The visibility tables hold a guard's cone only for the (waypoint, heading) pairs the patrol can
             reach; these traces pin that set against the reference itself.
* shape*   - grids that are not the usual squares (5x4, 7x12, 33x36, 40x64, 64x8, 6x16) with get_state_tensor
             (environment.py:347-374) recorded after EVERY step: the fused tick kernel writes that tensor with its own
             index arithmetic (channels shorter than a warp trip, rows of two words, two rows per lane).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the reference on sys.path, imports it)

OUT = os.path.join(HERE, "heist_golden_r3.npz")


def patrol_layout(rng, kind):
    R = C = 20
    walls = [(int(rng.integers(2, R - 2)), int(rng.integers(2, C - 2))) for _ in range(int(rng.integers(2, 7)))]
    cams = [{"row": int(rng.integers(3, R - 3)), "col": int(rng.integers(3, C - 3)), "fov_angle": float(np.float32(rng.uniform(40, 100))),
             "heading": float(np.float32(rng.uniform(0, 360))), "rotation_speed": float(np.float32(rng.uniform(5, 35))), "vision_range": 5}]
    guards = []
    n_guards = {"ring": 2, "stride": 3, "still": 2, "nomove": 2, "crowd": 4}[kind]
    for g in range(n_guards):
        r0, c0 = int(rng.integers(3, 9)), int(rng.integers(3, 9))   # near the Solver's start: cones reach its first tiles
        ring = mg.patrol(r0, c0, R, C)
        if kind == "ring":
            path, speed = ring, [1, -1][g % 2]
        elif kind == "stride":
            path, speed = ring[:int(rng.integers(5, 9))], [2, 3, 5][g % 3]
        elif kind == "still":
            path, speed = (ring[:1], 1) if g == 0 else (ring[:4], 8)            # one waypoint / stride = 2 x length
        elif kind == "nomove":
            path = [ring[0], ring[0], ring[1], ring[2], ring[2], ring[2], ring[3]][:int(rng.integers(4, 8))]
            speed = [1, 2][g % 2]
        else:
            path, speed = ring[:int(rng.integers(2, 9))], int(rng.choice([1, 1, 2, -1, 3]))
        guards.append({"patrol_path": path, "speed": int(speed), "vision_range": int(rng.choice([3, 4, 4])),
                       "fov_angle": float(rng.choice([90.0, 90.0, 60.0, 120.0]))})
    return walls, cams, guards, 1000


def shape_layout(rng, R, C):
    walls = [(int(rng.integers(1, R - 1)), int(rng.integers(1, C - 1))) for _ in range(int(rng.integers(0, 4)))]
    walls = [w for w in walls if w not in ((1, 1), (R - 2, C - 2))]
    cams = [{"row": int(rng.integers(1, R - 1)), "col": int(rng.integers(1, C - 1)), "fov_angle": float(np.float32(rng.uniform(40, 110))),
             "heading": float(np.float32(rng.uniform(0, 360))), "rotation_speed": float(np.float32(rng.uniform(5, 35))),
             "vision_range": int(rng.integers(2, 7))} for _ in range(int(rng.integers(1, 3)))]
    guards = [{"patrol_path": mg.patrol(int(rng.integers(1, R - 1)), int(rng.integers(1, C - 1)), R, C), "speed": 1,
               "vision_range": 3, "fov_angle": 90.0}]
    return walls, cams, guards, 1000


def main():
    store, meta = {}, {"traces": [], "numpy": np.__version__}
    rng = np.random.default_rng(20261020)
    cases = []
    for kind in ("ring", "stride", "still", "nomove", "crowd"):
        for k in range(2):
            ms = [7, 11, 13, 9, 10][len(cases) % 5]
            cases.append((f"patrol_{kind}_{k}", 20, 20, ms, patrol_layout(rng, kind), rng.integers(0, 5, 150).astype(np.int8)))
    for R, C in ((5, 4), (7, 12), (33, 36), (40, 64), (64, 8), (6, 16)):
        cases.append((f"shape_{R}x{C}", R, C, 9, shape_layout(rng, R, C), mg.biased_actions(rng, 48)))
    for name, R, C, ms, layout, actions in cases:
        rec = mg.run_trace(R, C, ms, layout, actions, want_state_every=1 if name.startswith("shape") else 17)
        walls, cams, guards, budget = layout
        meta["traces"].append({"name": name, "R": R, "C": C, "max_steps": ms, "budget": budget,
                               "walls": [list(map(int, w)) for w in walls], "cameras": cams,
                               "guards": [{**g, "patrol_path": [list(map(int, p)) for p in g["patrol_path"]]} for g in guards],
                               "valid": rec["valid"], "spent": int(rec["spent"]), "n_placed": list(map(int, rec["n_placed"]))})
        store[f"{name}/actions"] = actions
        for key in ["grid", "vis0", "reward", "done", "status", "pos", "tick", "vis", "vis_post", "cam_heading",
                    "guard_idx", "guard_heading", "state_t", "state", "obs_vec"]:
            store[f"{name}/{key}"] = rec[key]
        print("trace", name, "valid", rec["valid"], "episodes ended", int(np.sum(rec["done"])),
              "detected", int(np.sum(rec["status"] == mg.STATUS["detected"])), flush=True)
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(meta["traces"]), "traces")


if __name__ == "__main__":
    main()
