"""Generate tests/golden/preprocess.npz by running the reference's own pre-processing functions
(/root/reference/utils/utils.py: resize_image, preprocess_input, preprocess_input_radar; achelous.py:224-246 with
sklearn.preprocessing.normalize) on seeded inputs, with the Pillow / numpy / sklearn installed in this container.

    python tests/golden/make_preprocess_golden.py

Inputs are regenerated from the seeds by tests/test_preprocess.py (`make_case`), only outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np
from PIL import Image
from sklearn.preprocessing import normalize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_preprocess import IMAGE_CASES, make_image, make_points, make_radar  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/utils/utils.py")
ref_utils = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_utils)


def main():
    out = {}
    for name, (ih, iw, h, w, letterbox, seed) in IMAGE_CASES.items():
        img = make_image(ih, iw, seed)
        image = Image.fromarray(img)
        data = ref_utils.resize_image(image, (w, h), letterbox)                                   # achelous.py:200
        x = np.expand_dims(np.transpose(ref_utils.preprocess_input(np.array(data, dtype="float32")), (2, 0, 1)), 0)   # :205
        out["image." + name] = x.astype(np.float32)
    for name, dtype in (("radar.f64", np.float64), ("radar.f32", np.float32)):
        out[name] = ref_utils.preprocess_input_radar(make_radar(dtype)).astype(np.float32)        # achelous.py:212
    feats, idx = make_points()
    out["points"] = np.ascontiguousarray(np.array(normalize(X=feats[idx], axis=0), dtype=np.float32).T)[None]   # achelous.py:240-243
    path = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
