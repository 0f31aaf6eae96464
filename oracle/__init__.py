"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the per-frame pixel hot path of ayf7/cuauv-vision-pipeline.  Nothing under this
package is imported by the product (`cuauv_vision_pipeline_b200`); only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may use it, and
only as the checker / the reported CPU baseline.

Pieces
------
ref_balance      ctypes call into oracle/_ref/libauv-color-balance-ref.so, i.e. the reference's own
                 utils/color_correction/color_balance.cpp compiled unmodified (oracle/Makefile),
                 invoked with the exact marshalling of modules/color_balance.py:93-110.
color_balance_np numpy + cv2 restatement of process_frame (default and P1 flags), validated
                 byte-for-byte against ref_balance (tests/test_oracle_ref.py).
cv_ops           the literal cv2 calls the reference makes (utils/color.py, utils/transform.py,
                 utils/feature.py, modules/bins.py, modules/preprocessor.py).  OpenCV is a third-party
                 dependency of the reference (configure.py:27-33, `pkg-config opencv4`, unpinned);
                 parity is pinned to the installed cv2 4.13.0.
spec_np          independent numpy restatements of OpenCV's published 8-bit algorithms (BGR2HSV,
                 HSV2BGR, BGR2LAB, BGR2GRAY, BGR2YCrCb, BGR2HLS, INTER_LINEAR resize); they pin the
                 arithmetic the CUDA kernels implement and are themselves checked against cv2.
ccl              declared oracle for labelling + raster moments: cv2
                 connectedComponentsWithStats(8) canonicalised by first-pixel raster order + exact
                 integer moments.
letterbox        restatement of the Ultralytics LetterBox input transform (parity unpinned: the
                 package is neither vendored nor installed).
synth            seeded synthetic frame / mask generators.

Parity status: the reference repository contains no tests, golden vectors or fixtures
(build.ninja:72-73 are empty phony targets), so every row is pinned by outputs of the reference
itself run here: the compiled color_balance.cpp and the cv2 4.13.0 calls.  The YOLO letterbox
row alone is "parity unpinned".
"""
