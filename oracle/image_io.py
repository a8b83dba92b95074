"""CPU restatement of the reference's image writer arithmetic (TEST INFRASTRUCTURE ONLY: imported by tests/, smoke()
and bench.py's cpu_baseline leg, never by the product path).

visualization.save_images (visualization.py:59-88): `tensor = (img_tensor.clone() + 1) * 0.5 * 255` (three separate
f32 roundings) -> `.cpu().clamp(0, 255)` -> `.numpy().astype("uint8")` (truncation toward zero) -> HWC for 3-channel
images, squeeze for 1-channel ones -> PIL PNG encoder (lossless, not restated).

Pinned: tests/golden/image_u8.npz holds the pixel arrays read back from PNGs written by the UNMODIFIED reference
function (oracle/make_golden_image.py); tests/test_oracle_cpu.py checks this restatement against them bit for bit.
"""
import numpy as np


def image_to_u8(x):
    """x: float32 array [B,C,H,W] -> uint8 [B,H,W,C] (visualization.py:73-80)."""
    x = np.asarray(x, dtype=np.float32)
    t = (x + np.float32(1)) * np.float32(0.5) * np.float32(255)
    t = np.clip(t, np.float32(0), np.float32(255))
    return np.ascontiguousarray(t.astype(np.uint8).transpose(0, 2, 3, 1))


def synth_images(seed, B=3, C=3, H=32, W=24):
    """Seeded images that cross both clamp edges and sit on rounding boundaries ((k/255)*2-1 exactly and +-1 ulp)."""
    rng = np.random.RandomState(seed)
    x = (rng.rand(B, C, H, W).astype(np.float32) * np.float32(2.6) - np.float32(1.3))
    k = rng.randint(0, 256, size=(H, W)).astype(np.float32)
    edge = k / np.float32(255) * np.float32(2) - np.float32(1)
    x[0, 0] = edge
    x[0, C - 1] = np.nextafter(edge, np.float32(-2))
    x[B - 1, 0] = np.nextafter(edge, np.float32(2))
    return x
