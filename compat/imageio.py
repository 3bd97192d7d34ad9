"""imageio.imread / imwrite over PIL — what the reference's Blender loader (src/data/loaders/load_blender.py:43) and
eval_nerf.py --save-images (src/eval_nerf.py:81-98) use."""
import numpy as np


def imread(path, *a, **k):
    from PIL import Image
    return np.asarray(Image.open(str(path)))


def imwrite(path, arr):
    from PIL import Image
    Image.fromarray(np.asarray(arr)).save(str(path))
