"""libavcodec (through cv2's FFmpeg backend) as an independent H.264 decoder: Annex-B bytes -> luma planes.

cv2 hands back the decoder's luma plane untouched when CAP_PROP_CONVERT_RGB is 0 (the chroma planes are not
reachable that way, and the BGR path converts colours), so the cross-check is on luma. TEST TOOLING ONLY."""
from __future__ import annotations

import os
import tempfile

import numpy as np


def available() -> bool:
    try:
        import cv2
        return "FFMPEG" in [cv2.videoio_registry.getBackendName(b) for b in cv2.videoio_registry.getStreamBackends()]
    except Exception:
        return False


def decode_luma(stream: bytes, n_frames: int, width: int, height: int) -> np.ndarray:
    """-> uint8 [n_frames, height, width]; raises if libavcodec does not return every picture."""
    import cv2
    os.environ.setdefault("OPENCV_FFMPEG_LOGLEVEL", "16")
    with tempfile.NamedTemporaryFile(suffix=".h264", delete=False) as f:
        f.write(stream)
        path = f.name
    try:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
        if not cap.isOpened():
            raise RuntimeError("libavcodec cannot open the stream")
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
        out = np.empty((n_frames, height, width), np.uint8)
        for f_idx in range(n_frames):
            ok, fr = cap.read()
            if not ok or fr is None or fr.size < width * height:
                raise RuntimeError(f"libavcodec returned {f_idx} of {n_frames} pictures")
            out[f_idx] = np.asarray(fr).reshape(-1)[:width * height].reshape(height, width)
        cap.release()
        return out
    finally:
        os.unlink(path)
