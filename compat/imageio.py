"""Minimal imageio.imwrite over PIL for eval_nerf.py --save-images."""
def imwrite(path, arr):
    from PIL import Image
    Image.fromarray(arr).save(path)
