"""Generates tests/golden/cv2_projection.json: pixel projections by OpenCV (an implementation independent of both the
reference and this repo) for the three camera models whose equations Calico shares with OpenCV:
  OpenCv5        = cv2.projectPoints with distCoeffs (k1,k2,p1,p2,k3)           (camera_models.h:105-141)
  OpenCv8        = cv2.projectPoints with (k1,k2,p1,p2,k3,k4,k5,k6)             (camera_models.h:257-298)
  KannalaBrandt  = cv2.fisheye.projectPoints with (k1,k2,k3,k4)                 (camera_models.h:420-462)
Intrinsics are the reference's own test constants (camera_models_test.cpp:107-108,128-130,153-154). Run once; the JSON is
committed so the tests need neither cv2 nor the reference."""
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
pts = np.concatenate([rng.uniform(-0.6, 0.6, size=(40, 2)), rng.uniform(0.6, 2.0, size=(40, 1))], axis=1)
out = {"points": pts.tolist(), "models": {}}
f, cx, cy = 785.0, 640.0, 400.0
K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
z3 = np.zeros(3)
d5 = np.array([-3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2])
d8 = np.array([-3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2, 1.225e-1, -5.26e-2, 8.58e-3])
d4 = np.array([-3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4])
px5, _ = cv2.projectPoints(pts.reshape(-1, 1, 3), z3, z3, K, d5)
px8, _ = cv2.projectPoints(pts.reshape(-1, 1, 3), z3, z3, K, d8)
pxk, _ = cv2.fisheye.projectPoints(pts.reshape(-1, 1, 3), z3, z3, K, d4)
out["models"]["1"] = {"intrinsics": [f, cx, cy, *d5.tolist()], "pixels": px5.reshape(-1, 2).tolist()}
out["models"]["2"] = {"intrinsics": [f, cx, cy, *d8.tolist()], "pixels": px8.reshape(-1, 2).tolist()}
out["models"]["3"] = {"intrinsics": [f, cx, cy, *d4.tolist()], "pixels": pxk.reshape(-1, 2).tolist()}
with open(os.path.join(HERE, "cv2_projection.json"), "w") as fh:
    json.dump(out, fh)
print("wrote", len(pts), "points x 3 models")
