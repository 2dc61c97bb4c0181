"""Row a1 robustness: malformed / hostile JPEG bytes must be REJECTED (RuntimeError, like libjpeg's error_exit ->
RuntimeError in the reference, dct_manip.cpp:24-41) and never touch memory outside the decoder's tables.

The decoder source is rebuilt with AddressSanitizer into a stand-alone harness (tests/harness/jpeg_asan_main.cpp); every
case below is run through rgbnm_jpeg_info_from_memory / rgbnm_jpeg_read_coefficients / rgbnm_jpeg_decode_batch under it."""
import io
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
from PIL import Image

from rgb_no_more_b200 import dct_manip as dm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    out = tmp_path_factory.mktemp("asan") / "jpeg_asan"
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
           "-pthread", os.path.join(ROOT, "tests", "harness", "jpeg_asan_main.cpp"),
           os.path.join(ROOT, "rgb_no_more_b200", "csrc", "jpeg_codec.cpp"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ASan build unavailable: " + r.stderr[-300:])
    return str(out)


def _valid_jpeg(size=512, quality=75, seed=3) -> bytes:
    rng = np.random.default_rng(seed)
    img = Image.fromarray(rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)).resize((size, size), Image.BICUBIC)
    b = io.BytesIO()
    img.save(b, "JPEG", quality=quality, subsampling=2)
    return b.getvalue()


def _seg(marker: int, body: bytes) -> bytes:
    return bytes([0xFF, marker]) + struct.pack(">H", len(body) + 2) + body


def _segments(buf: bytes):
    """[(marker, start, end)] of the header segments up to and including SOS."""
    pos, out = 2, []
    while pos + 4 <= len(buf):
        assert buf[pos] == 0xFF
        m = buf[pos + 1]
        ln = struct.unpack(">H", buf[pos + 2:pos + 4])[0]
        out.append((m, pos, pos + 2 + ln))
        pos += 2 + ln
        if m == 0xDA:
            break
    return out


def _hostile_cases():
    good = _valid_jpeg()
    segs = _segments(good)
    cases = {}
    # DHT whose counts are not a prefix code -> canonical code runs past look[512]
    bits = bytes([0, 200] + [0] * 14)
    cases["dht_overfull_len2"] = b"\xff\xd8" + _seg(0xC4, bytes([0x00]) + bits + bytes(200))
    cases["dht_overfull_len1"] = b"\xff\xd8" + _seg(0xC4, bytes([0x10]) + bytes([3] + [0] * 15) + bytes(3))
    cases["dht_overfull_len9"] = b"\xff\xd8" + _seg(0xC4, bytes([0x00]) + bytes([0] * 8 + [255] + [0] * 7) + bytes(255))
    cases["dht_truncated_counts"] = b"\xff\xd8" + _seg(0xC4, bytes([0x00, 1, 2, 3]))
    cases["dqt_truncated"] = b"\xff\xd8" + _seg(0xDB, bytes([0x00]) + bytes(10))
    cases["dqt_16bit_truncated"] = b"\xff\xd8" + _seg(0xDB, bytes([0x10]) + bytes(64))
    cases["sof_short"] = b"\xff\xd8" + _seg(0xC0, bytes([8, 0, 16, 0, 16, 3, 1]))
    cases["sof_empty"] = b"\xff\xd8" + _seg(0xC0, b"")
    cases["sos_short"] = good[:segs[-1][1]] + _seg(0xDA, bytes([3, 1]))
    cases["dri_short"] = b"\xff\xd8" + _seg(0xDD, b"")

    def patch(marker, offset, value, nth=0):
        k = [s for s in segs if s[0] == marker][nth]
        b = bytearray(good)
        b[k[1] + 4 + offset] = value
        return bytes(b)
    cases["sof_tq_255"] = patch(0xC0, 8, 255)                 # comp[0].tq indexes qt[4][64]
    cases["sof_tq_3_absent"] = patch(0xC0, 8, 3)
    cases["sof_zero_width"] = patch(0xC0, 3, 0)[:0] + bytes(bytearray(patch(0xC0, 3, 0)))  # high byte of width
    b = bytearray(good)
    k = [s for s in segs if s[0] == 0xC0][0]
    b[k[1] + 4 + 1:k[1] + 4 + 5] = bytes(4)                   # height = width = 0
    cases["sof_zero_size"] = bytes(b)
    cases["sos_td_15"] = patch(0xDA, 2, 0xF0)
    cases["sos_ta_15"] = patch(0xDA, 2, 0x0F)
    cases["sos_tables_absent"] = patch(0xDA, 2, 0x22)
    cases["sof_sampling_0"] = patch(0xC0, 7, 0x00)
    cases["truncated_scan"] = good[:segs[-1][2] + 100]
    cases["truncated_header"] = good[:segs[1][1] + 7]
    cases["empty"] = b""
    cases["soi_only"] = b"\xff\xd8"
    cases["not_jpeg"] = b"\x89PNG\r\n\x1a\n" + bytes(64)
    # seeded byte flips over the headers and the start of the scan
    rng = np.random.default_rng(11997733)
    hdr_end = segs[-1][2]
    for i in range(48):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(2, hdr_end + 64))] = int(rng.integers(0, 256))
        cases[f"flip_{i}"] = bytes(b)
    for i in range(8):
        cases[f"cut_{i}"] = good[:int(rng.integers(2, hdr_end + 8))]
    return good, cases


def test_hostile_inputs_under_asan(harness, tmp_path):
    good, cases = _hostile_cases()
    names = ["good"] + sorted(cases)
    for n in names:
        (tmp_path / n).write_bytes(good if n == "good" else cases[n])
    r = subprocess.run([harness] + [str(tmp_path / n) for n in names], capture_output=True, text=True,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0"))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(names)
    assert lines[0] == "0 0 0"                                 # the unmodified file decodes through all three entry points
    rc = dict(zip(names, (tuple(int(x) for x in ln.split()) for ln in lines)))
    for n in ("dht_overfull_len2", "dht_overfull_len1", "dht_overfull_len9", "dht_truncated_counts", "dqt_truncated",
              "dqt_16bit_truncated", "sof_short", "sof_empty", "sos_short", "sof_tq_255", "sof_tq_3_absent", "sof_zero_size",
              "sos_td_15", "sos_ta_15", "sos_tables_absent", "sof_sampling_0", "empty", "soi_only", "not_jpeg", "truncated_header"):
        assert rc[n][1] != 0 and rc[n][2] != 0, (n, rc[n])


def test_hostile_inputs_raise_runtime_error():
    """Same contract through the Python module (boundary B1): RuntimeError, no partial result."""
    _, cases = _hostile_cases()
    for n in ("dht_overfull_len2", "sof_tq_255", "sos_td_15", "sof_zero_size", "empty"):
        with pytest.raises(RuntimeError):
            dm.read_coefficients_from_bytes(cases[n])
