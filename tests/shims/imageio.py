"""TEST SHIM (not product code): minimal `imageio` stand-in for images without the real package, so that the reference's
UNMODIFIED scripts and dataset adapters (which `import imageio`) can run in tests.  PNG/JPG through OpenCV; videos are
written as a .npy next to the requested path."""
import cv2
import numpy as np


def imread(path, *a, **k):
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    return img


def imwrite(path, img, *a, **k):
    img = np.asarray(img)
    if img.ndim == 3 and img.shape[2] >= 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    if not cv2.imwrite(str(path), img):
        raise IOError(path)


imsave = imwrite


def mimwrite(path, frames, *a, **k):
    np.save(str(path) + ".npy", np.asarray(frames))


mimsave = mimwrite
