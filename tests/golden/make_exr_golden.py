#!/usr/bin/env python3
"""Writes tests/golden/exr/*.exr with OpenCV, i.e. with the OpenEXR library itself -- an encoder
independent of this repository's reader and of the test-suite's own minimal writer -- plus the
source pixels (exr_source.npz).  test_host_parser.py::test_exr_files_written_by_openexr reads them.

    OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/make_exr_golden.py
"""
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exr")


def main():
    os.makedirs(HERE, exist_ok=True)
    yy, xx = np.mgrid[0:45, 0:61]
    rng = np.random.RandomState(11)
    # smooth ramps (wavelet / predictor friendly), a saturated HDR patch, a noisy band, exact zeros
    rgb = np.stack([np.sin(xx / 9.0) + 1.5 + yy / 30.0, np.cos(yy / 5.0) * 0.5 + 0.7, (xx * yy) / 900.0], axis=2).astype(np.float32)
    rgb[8:14, 20:40] = (900.0, 650.0, 410.0)
    rgb[30:36] += rng.rand(6, 61, 3).astype(np.float32) * 0.01
    rgb[40:, :10] = 0.0
    np.savez_compressed(os.path.join(HERE, "exr_source.npz"), rgb=rgb)
    codecs = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS,
              "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP, "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}
    types = {"half": cv2.IMWRITE_EXR_TYPE_HALF, "float": cv2.IMWRITE_EXR_TYPE_FLOAT}
    for cname, tname in (("piz", "half"), ("piz", "float"), ("zip", "half"), ("zips", "float"), ("rle", "half"), ("none", "float")):
        path = os.path.join(HERE, "%s_%s.exr" % (cname, tname))
        assert cv2.imwrite(path, rgb[:, :, ::-1], [cv2.IMWRITE_EXR_TYPE, types[tname], cv2.IMWRITE_EXR_COMPRESSION, codecs[cname]])   # OpenCV is BGR
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
