"""Golden vectors for the 3D-mask projection front end, generated from the reference's OWN functions.

scripts/project_3d_masks.py imports pytorch3d at module level (absent here), so the three pure-numpy functions the
projection's geometry rests on -- generate_predicted_grid (:108-131), grid2world (:69-74), grid_pts_coord (:77-88) -- are
extracted from the reference file by name (ast) and executed unmodified.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_project.py        ->  tests/golden/ref_project.npz
"""
import ast
import os

import numpy as np

REF = "/root/reference/instance_nerf/scripts/project_3d_masks.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_project.npz")


def load_functions(names):
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return [ns[n] for n in names]


def main():
    gen_grid, grid2world, grid_pts_coord = load_functions(["generate_predicted_grid", "grid2world", "grid_pts_coord"])
    rng = np.random.default_rng(0)
    shape, M = (12, 9, 14), 5
    # overlapping box-shaped masks with distinct scores, as the 3D detector's output looks (masks [M, X, Y, Z] -> transposed :184)
    masks = np.zeros((M,) + shape, dtype=bool)
    for m in range(M):
        lo = [rng.integers(0, s - 3) for s in shape]
        hi = [min(s, l + rng.integers(2, 7)) for s, l in zip(shape, lo)]
        masks[m, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
    scores = rng.random(M).astype(np.float64)
    masks_t = masks.transpose([1, 2, 3, 0])
    labels = gen_grid(masks_t.copy(), scores)                    # flat [X*Y*Z]
    room_bbox = np.array([[-3.0, -1.4, -2.5], [3.1, 1.5, 2.4]])
    pts = grid_pts_coord(masks_t[..., 0], room_bbox=room_bbox)  # [X*Y*Z, 3]
    np.savez_compressed(OUT, masks=masks, scores=scores, labels=np.asarray(labels).reshape(shape), room_bbox=room_bbox, points=pts)
    print("wrote", OUT, "labels:", np.unique(np.asarray(labels)))


if __name__ == "__main__":
    main()
