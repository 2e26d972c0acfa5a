"""CPU check of the single-pass 2-D scan's inter-tile protocol (shared_b200/csrc/scan2d_core.cuh -- the ticket order, the
descriptor words and the look-back array_scan2d.cu's kernel runs): host threads stand in for CTAs (tests/emu/scan2d_emu.cpp).
The result must equal the serial summed-area table bit for bit on integer-valued data (what rdOp over both dimensions,
DimensionOps.java:411-435, and createIntegralImage, NativeImageKernel.cpp:32-118, produce), with any number of workers down
to ONE -- a tile never waits for a ticket that has not been handed out -- and without a data race under ThreadSanitizer."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("no g++ / make")
    res = subprocess.run(["make", "-s", "-k", "-C", EMU, "all"], capture_output=True, text=True)
    fast = os.path.join(EMU, "_build", "scan2d_emu_fast")
    tsan = os.path.join(EMU, "_build", "scan2d_emu_tsan")
    if not os.path.exists(fast):
        pytest.fail("emulation build failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    return {"fast": fast, "tsan": tsan if os.path.exists(tsan) else None}


def scan(exe, rows, cols, workers, seed=1, border=0, inplace=0, timeout=120):
    return subprocess.run([exe, "scan"] + [str(v) for v in (rows, cols, workers, seed, border, inplace)], capture_output=True,
                          text=True, timeout=timeout)


def test_ticket_order_is_a_bijection_with_predecessors_first(emu):
    # every tile grid up to 130 x 130 (8192 x 8192 is 128 x 128 tiles): each tile exactly once, left / upper neighbour earlier
    res = subprocess.run([emu["fast"], "tickets", "130", "130"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-500:]


@pytest.mark.parametrize("rows,cols", [(64, 64), (65, 63), (300, 450), (1, 1000), (1000, 3), (129, 129), (640, 130)])
def test_scan_matches_serial_table_for_any_worker_count(emu, rows, cols):
    for workers in (1, 2, 5, 16):  # one worker: the hand-out order alone must guarantee progress
        res = scan(emu["fast"], rows, cols, workers, seed=rows * 31 + cols)
        assert res.returncode == 0, (rows, cols, workers, res.stderr[-500:])


def test_integral_image_border_and_in_place(emu):
    for rows, cols in ((301, 451), (65, 65), (2, 700)):
        res = scan(emu["fast"], rows, cols, 4, seed=5, border=1)
        assert res.returncode == 0, (rows, cols, res.stderr[-500:])
    res = scan(emu["fast"], 200, 330, 4, seed=6, inplace=1)
    assert res.returncode == 0, res.stderr[-500:]


def test_protocol_has_no_data_race(emu):
    if emu["tsan"] is None:
        pytest.skip("no ThreadSanitizer runtime")
    for rows, cols, workers, border in ((200, 330, 4, 0), (130, 260, 8, 1), (70, 700, 3, 0)):
        res = scan(emu["tsan"], rows, cols, workers, seed=3, border=border, timeout=600)
        assert res.returncode == 0 and "ThreadSanitizer" not in res.stderr, (rows, cols, res.stderr[-1500:])
