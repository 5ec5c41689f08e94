"""libavcodec (through cv2's FFmpeg backend) as an independent H.264 decoder: Annex-B bytes -> luma planes.

cv2 hands back the decoder's luma plane untouched when CAP_PROP_CONVERT_RGB is 0; the chroma planes are not reachable
that way. They are compared through the BGR path instead: libavformat also reads YUV4MPEG2, so the candidate pictures are
written as a .y4m file and go through the very same swscale yuv420p -> bgr24 conversion as the decoded stream; equal BGR
pictures pin the chroma planes (a one-LSB change of a chroma sample moves B or R by 1.6 - 2 and is seen unless every pixel
it touches is clipped). TEST TOOLING ONLY."""
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


def decode_bgr(data: bytes, n_frames: int, suffix: str = ".h264") -> np.ndarray:
    """-> uint8 [n_frames, height, width, 3]: what cv2 returns by default (libavcodec / rawvideo + swscale to bgr24)."""
    import cv2
    os.environ.setdefault("OPENCV_FFMPEG_LOGLEVEL", "16")
    with tempfile.NamedTemporaryFile(suffix=suffix, delete=False) as f:
        f.write(data)
        path = f.name
    try:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
        if not cap.isOpened():
            raise RuntimeError("libavformat cannot open the file")
        out = []
        for f_idx in range(n_frames):
            ok, fr = cap.read()
            if not ok or fr is None:
                raise RuntimeError(f"libavcodec returned {f_idx} of {n_frames} pictures")
            out.append(np.asarray(fr).copy())
        cap.release()
        return np.stack(out)
    finally:
        os.unlink(path)


def y4m(frames: np.ndarray, width: int, height: int) -> bytes:
    """YUV4MPEG2 file of planar 4:2:0 pictures (uint8 [n, width*height*3/2], Y | Cb | Cr: the reconstruction's layout)."""
    out = bytearray(f"YUV4MPEG2 W{width} H{height} F25:1 Ip A1:1 C420mpeg2\n".encode())
    for f in np.asarray(frames, np.uint8).reshape(len(frames), -1):
        assert f.size == width * height * 3 // 2
        out += b"FRAME\n" + f.tobytes()
    return bytes(out)


def bgr_of_pictures(frames: np.ndarray, width: int, height: int) -> np.ndarray:
    """The BGR pictures cv2 makes of planar 4:2:0 pictures: same conversion as decode_bgr applies to a decoded stream."""
    return decode_bgr(y4m(frames, width, height), len(frames), ".y4m")
