"""TEST INFRASTRUCTURE -- the reference's post-render image expressions, evaluated with numpy float32 on the CPU
(IEEE float32 elementwise operations, the same ones torch executes):

    street_gaussian_renderer.py:340   rgb = rgb + sky_color * (1 - acc)
    street_gaussian_renderer.py:345   rgb = clamp(rgb, 0, 1)                      (cfg.mode != 'train')
    simulator.py:314                  (rgb.cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)

Only tests/ import this."""
import numpy as np


def compose_rgb8(rgb, acc=None, sky=None):
    rgb = np.asarray(rgb, dtype=np.float32)
    if sky is not None:
        t = np.asarray(sky, dtype=np.float32) * (np.float32(1.0) - np.asarray(acc, dtype=np.float32))
        rgb = rgb + t
    rgb = np.clip(rgb, np.float32(0.0), np.float32(1.0))
    return (rgb.transpose(1, 2, 0) * 255).astype(np.uint8), rgb
