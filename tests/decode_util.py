"""Independent HEVC decoder for the conformance tests: libavcodec through OpenCV's FFmpeg backend (TEST
INFRASTRUCTURE ONLY).  Returns the decoded luma plane, or None when this OpenCV build cannot decode HEVC."""
import os
import tempfile

import numpy as np


def decode_luma(stream, shape):
    try:
        import cv2
    except ImportError:
        return None
    fd, path = tempfile.mkstemp(suffix=".h265")
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(stream)
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
        ok, frame = cap.read()
        cap.release()
    finally:
        os.unlink(path)
    if not ok:
        return None
    flat = np.asarray(frame).reshape(-1)
    n = shape[0] * shape[1]
    if flat.size < n:
        return None
    return flat[:n].reshape(shape).copy()   # the frame is 4:2:0: the luma plane comes first
