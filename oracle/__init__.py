"""CPU oracle for the lane-detection hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The
product package never imports it.

Two layers:

* :mod:`oracle.stages` -- ctypes wrappers over ``lane_oracle.c`` (a C restatement of the
  OpenCV/NumPy arithmetic behind ``/root/reference/src/perception/lane_detector.py:66-101``)
  plus NumPy restatements of the fit / EMA / offset tail (``:105-176``, ``:253-272``).
  It exposes the intermediates cv2 hides.
* :mod:`oracle.cv2_pipeline` -- the reference's call sequence restated on top of the real
  ``cv2`` / ``numpy`` (the same third-party binaries the reference runs on), used as the
  final arbiter and as the CPU baseline.

Parity pin: the reference ships no tests or golden vectors.  ``tests/test_oracle_vs_cv2.py``
pins the C restatement against cv2 4.13.0 / numpy 2.3.5 in this image, and
``tests/golden/`` holds outputs of the unmodified reference ``LaneDetector`` produced by
``tests/golden/make_golden.py`` (run where ``/root/reference`` is mounted).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "lane_oracle.c")
_LIB = os.path.join(_HERE, "_build", "liblane_oracle.so")


def build(force: bool = False) -> str:
    """Compile lane_oracle.c with gcc (no FMA contraction) and return the .so path."""
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fvisibility=hidden", "-shared", "-fPIC",
             "-o", _LIB, _SRC, "-lm"])
    return _LIB


_EMU_SRC = os.path.join(os.path.dirname(_HERE), "tests", "native", "draw_emulator.cpp")
_EMU_HDR = os.path.join(os.path.dirname(_HERE), "multimodal_autonomous_driving_perception_and_planning_b200", "csrc",
                        "draw_prims.h")
_EMU_LIB = os.path.join(_HERE, "_build", "libdraw_emulator.so")


def build_draw_emulator(force: bool = False) -> str:
    """Compile tests/native/draw_emulator.cpp (the CPU replay of K7's device primitives: test infrastructure that lets
    the drawing path's HOST half be checked against cv2 without a GPU) and return the .so path."""
    os.makedirs(os.path.dirname(_EMU_LIB), exist_ok=True)
    newest = max(os.path.getmtime(_EMU_SRC), os.path.getmtime(_EMU_HDR))
    if force or not os.path.exists(_EMU_LIB) or os.path.getmtime(_EMU_LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-shared", "-fPIC", "-o", _EMU_LIB, _EMU_SRC])
    return _EMU_LIB


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
