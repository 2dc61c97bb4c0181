"""Row a1 / boundary B1 on the CPU: the from-scratch baseline-JPEG coefficient reader behind the C-ABI
(rgbnm_jpeg_*; replaces dct_manip.read_coefficients, /root/reference/dct_manip/dct_manip.cpp:78-178).

The reference has no tests or fixtures at this boundary and its libjpeg is unpinned, so the decoder is
pinned by the JPEG standard itself: (i) bit-exact round trips through the coefficient writer, (ii) files encoded by an
independent encoder (PIL / libjpeg-turbo) whose decoded coefficients, pushed through a float IDCT, reproduce that
library's own decoded pixels -- any Huffman / marker / ordering slip produces gross errors."""
import io

import numpy as np
import pytest
import torch
from PIL import Image
from scipy.fft import idctn

from rgb_no_more_b200 import dct_manip as dm
from rgb_no_more_b200 import synth


def _blocks_to_plane(coef: torch.Tensor, q: torch.Tensor) -> np.ndarray:
    """(hb, wb, 8, 8) quantised blocks -> (8 hb, 8 wb) float pixels (level shift included)."""
    deq = (coef.float() * q.float()).numpy()
    pix = idctn(deq, axes=(2, 3), norm="ortho") + 128.0
    hb, wb = coef.shape[:2]
    return pix.transpose(0, 2, 1, 3).reshape(hb * 8, wb * 8)


def _pil_ycc(buf: bytes) -> np.ndarray:
    im = Image.open(io.BytesIO(buf))
    im.draft("YCbCr", im.size)                       # libjpeg's own YCbCr output, no colour conversion
    assert im.mode == "YCbCr"
    return np.asarray(im).astype(np.float32)


@pytest.mark.parametrize("chroma", [(2, 2), (1, 1)])
@pytest.mark.parametrize("hw", [(64, 64), (72, 40)])
def test_roundtrip_through_the_coefficient_writer(chroma, hw):
    h, w = hw
    rng = np.random.default_rng(h * 7 + chroma[0])
    # like libjpeg's height_in_blocks / width_in_blocks (dct_manip.cpp:80-93) the planes cover ceil(size / 8) blocks of
    # the (down-sampled) component; blocks that only pad the last MCU are neither passed to the writer nor returned
    ch, cw = -(-h // chroma[1]), -(-w // chroma[0])
    yb, xb, cyb, cxb = -(-h // 8), -(-w // 8), -(-ch // 8), -(-cw // 8)
    y = torch.from_numpy(rng.integers(-1023, 1024, size=(1, yb, xb, 8, 8)).astype(np.int16))
    y[..., 0, 0] = torch.from_numpy(rng.integers(-1024, 1024, size=(1, yb, xb)).astype(np.int16))
    c = torch.from_numpy(rng.integers(-1023, 1024, size=(2, cyb, cxb, 8, 8)).astype(np.int16))
    q = torch.from_numpy(rng.integers(1, 256, size=(3, 8, 8)).astype(np.int16))
    buf = dm.write_coefficients(w, h, y, c, q, chroma=chroma)
    dims, q2, y2, c2 = dm.read_coefficients_from_bytes(buf)
    assert y2.shape == y.shape and c2.shape == c.shape
    assert torch.equal(y2, y) and torch.equal(c2, c) and torch.equal(q2, q)
    assert dims.tolist() == [[h, w], [ch, cw], [ch, cw]]
    # `dimensions` = downsampled component sizes (dct_manip.cpp:117-120)
    assert dims.dtype == torch.int32
    assert y2.dtype == torch.int16 and q2.dtype == torch.int16


def test_grayscale_has_no_chroma():
    rng = np.random.default_rng(5)
    y = torch.from_numpy(rng.integers(-500, 500, size=(1, 4, 4, 8, 8)).astype(np.int16))
    q = torch.from_numpy(rng.integers(1, 100, size=(1, 8, 8)).astype(np.int16))
    dims, q2, y2, c2 = dm.read_coefficients_from_bytes(dm.write_coefficients(32, 32, y, None, q))
    assert c2 is None and torch.equal(y2, y) and q2.shape == (1, 8, 8) and dims.shape == (1, 2)      # dct_manip.cpp:127-129


@pytest.mark.parametrize("quality", [50, 75, 100])
def test_pil_encoded_420_matches_libjpeg_pixels(quality):
    rng = np.random.default_rng(quality)
    low = rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)
    img = Image.fromarray(low).resize((512, 512), Image.BICUBIC)
    b = io.BytesIO()
    img.save(b, "JPEG", quality=quality, subsampling=2)              # 4:2:0 baseline, default Huffman tables (mp_scripts.py:74-81)
    buf = b.getvalue()
    dims, quant, y, c = dm.read_coefficients_from_bytes(buf)
    assert y.shape == (1, 64, 64, 8, 8) and c.shape == (2, 32, 32, 8, 8) and dims.tolist() == [[512, 512], [256, 256], [256, 256]]
    pil_q = Image.open(io.BytesIO(buf)).quantization
    assert quant[0].reshape(-1).tolist() == list(pil_q[0]) and quant[1].reshape(-1).tolist() == list(pil_q[1])   # natural order
    ycc = _pil_ycc(buf)
    ours = np.clip(np.round(_blocks_to_plane(y[0], quant[0])), 0, 255)
    assert np.abs(ours - ycc[..., 0]).max() <= 1.0                                                      # libjpeg's integer IDCT: <= 1 level
    # chroma: decode at scale 1/2 -- the output then has the chroma planes' own resolution, libjpeg does no up-sampling and
    # Cb / Cr are the plain 8x8 IDCT of our blocks
    half = Image.open(io.BytesIO(buf))
    half.draft("YCbCr", (256, 256))
    assert half.size == (256, 256) and half.mode == "YCbCr"
    hcc = np.asarray(half).astype(np.float32)
    for k in (0, 1):
        ours_c = np.clip(np.round(_blocks_to_plane(c[k], quant[1 + k])), 0, 255)
        assert np.abs(ours_c - hcc[..., 1 + k]).max() <= 1.0


def test_pil_encoded_444_matches_libjpeg_pixels_all_planes():
    img = Image.open(io.BytesIO(synth.synth_jpeg(1, 128))).convert("RGB")
    b = io.BytesIO()
    img.save(b, "JPEG", quality=90, subsampling=0)
    buf = b.getvalue()
    dims, quant, y, c = dm.read_coefficients_from_bytes(buf)
    assert c.shape == (2, 16, 16, 8, 8)
    ycc = _pil_ycc(buf)
    planes = [_blocks_to_plane(y[0], quant[0]), _blocks_to_plane(c[0], quant[1]), _blocks_to_plane(c[1], quant[2])]
    for k in range(3):
        assert np.abs(np.clip(np.round(planes[k]), 0, 255) - ycc[..., k]).max() <= 1.0


def test_batch_decode_equals_single_decode_and_reports_the_clamp():
    jpegs = synth.synth_jpeg_set(6)
    y, c, q, flags = dm.decode_batch(jpegs, 64, 64, nthreads=3)
    for i, buf in enumerate(jpegs):
        _, q1, y1, c1, fl = dm.read_coefficients_from_bytes(buf, return_clamp_flag=True)
        assert torch.equal(y[i].view(64, 64, 8, 8), y1[0]) and torch.equal(c[i].view(2, 32, 32, 8, 8), c1)
        assert torch.equal(q[i].view(3, 8, 8), q1) and bool(flags[i]) == fl
        yq = y1.int() * q1[0].int()
        cq = c1.int() * q1[1:3, None, None].int()
        live = bool(yq.min() < -1024 or yq.max() > 1016 or cq.min() < -1024 or cq.max() > 1016)
        assert fl == live                                            # datasets.py:288-290 clamp is live iff flagged
    # decoding again into the same staging buffers (a feeder ring slot) gives the same planes
    y0, c0, q0 = y.clone(), c.clone(), q.clone()
    y.zero_()
    y2, c2, q2, flags2 = dm.decode_batch(jpegs, 64, 64, nthreads=2, out=(y, c, q))
    assert y2.data_ptr() == y.data_ptr() and torch.equal(y2, y0) and torch.equal(c2, c0) and torch.equal(q2, q0)
    assert torch.equal(flags2, flags)


def test_errors_are_runtime_errors_like_the_pybind_module():
    with pytest.raises(RuntimeError):
        dm.read_coefficients("/nonexistent/file.jpg")                # dct_manip.cpp:155-159
    with pytest.raises(RuntimeError):
        dm.read_coefficients_from_bytes(b"not a jpeg at all")
    b = io.BytesIO()
    Image.open(io.BytesIO(synth.synth_jpeg(0, 64))).save(b, "JPEG", progressive=True)
    with pytest.raises(RuntimeError):
        dm.read_coefficients_from_bytes(b.getvalue())                # SOF2: outside the path, reported, not mis-decoded
    with pytest.raises(RuntimeError):
        dm.decode_batch([synth.synth_jpeg(0, 64)], 64, 64)           # wrong geometry for the batch layout


def _scan_bytes(buf: bytes) -> bytes:
    """Entropy-coded segment: everything after the SOS header up to the EOI marker."""
    pos = 2
    while True:
        assert buf[pos] == 0xFF
        m = buf[pos + 1]
        ln = (buf[pos + 2] << 8) | buf[pos + 3]
        pos += 2 + ln
        if m == 0xDA:
            break
    end = buf.rindex(b"\xff\xd9")
    return buf[pos:end]


@pytest.mark.parametrize("quality,subsampling,size", [(75, 2, (512, 512)), (50, 2, (512, 512)), (100, 2, (512, 512)),
                                                      (90, 0, (512, 512)), (75, 2, (200, 136)), (95, 0, (72, 40))])
def test_pil_file_reencodes_to_identical_scan_bytes(quality, subsampling, size):
    """Bit-exact pin of the Huffman decoder against libjpeg-turbo's ENCODER: PIL writes baseline
    files with the Annex-K tables; decoding such a file and writing the coefficients back with the same tables must
    reproduce its entropy-coded segment byte for byte.  A single wrong coefficient (+-1, wrong position, wrong
    DC prediction) changes the re-encoded bits, which the <= 1 level pixel comparison above cannot see at q = 1."""
    rng = np.random.default_rng(quality * 31 + subsampling)
    low = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    img = Image.fromarray(low).resize(size, Image.BICUBIC)
    noise = rng.integers(-12, 13, size=(size[1], size[0], 3))
    img = Image.fromarray(np.clip(np.asarray(img).astype(np.int16) + noise, 0, 255).astype(np.uint8))
    b = io.BytesIO()
    img.save(b, "JPEG", quality=quality, subsampling=subsampling)
    buf = b.getvalue()
    dims, quant, y, c = dm.read_coefficients_from_bytes(buf)
    chroma = (2, 2) if subsampling == 2 else (1, 1)
    again = dm.write_coefficients(size[0], size[1], y, c, quant, chroma=chroma)
    a, b2 = _scan_bytes(buf), _scan_bytes(again)
    assert len(a) == len(b2)
    assert a == b2
    # and an independent decoder (libjpeg-turbo through PIL) reads identical pixels from both files
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(again)).convert("RGB")), np.asarray(Image.open(io.BytesIO(buf)).convert("RGB")))


@pytest.mark.parametrize("quality,subsampling,size", [(75, 2, (512, 512)), (100, 2, (200, 136)), (30, 0, (97, 61))])
def test_restart_interval_streams_decode_to_the_same_coefficients(quality, subsampling, size):
    """The decoder has two scan paths: streams without restart markers go through the un-stuffed, branch-free bit reader
    (`decode_scan_fast`), DRI streams through the byte-wise one.  libjpeg writes the SAME coefficients with and without restart
    markers (only the DC predictions are reset), so both paths must return identical planes, tables and clamp flags."""
    rng = np.random.default_rng(quality + size[0])
    w, h = size
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 90 * np.sin(xx / 9.0 + c) * np.cos(yy / 13.0) + rng.normal(0, 18, (h, w)) for c in range(3)], -1)
    im = Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))
    plain, dri = io.BytesIO(), io.BytesIO()
    im.save(plain, "JPEG", quality=quality, subsampling=subsampling)
    im.save(dri, "JPEG", quality=quality, subsampling=subsampling, restart_marker_rows=1)
    assert b"\xff\xdd" in dri.getvalue() and b"\xff\xdd" not in plain.getvalue()
    a = dm.read_coefficients_from_bytes(plain.getvalue())
    b = dm.read_coefficients_from_bytes(dri.getvalue())
    for ta, tb in zip(a, b):
        assert (ta is None) == (tb is None)
        if ta is not None:
            assert torch.equal(ta, tb)


def test_plan_first_decode_stops_after_the_last_needed_block_row():
    """decode_batch(last_rows=...): the scan is abandoned after the MCU row holding the crop window's last luma block row;
    everything up to there equals the full decode, the block rows below keep the staging buffer's previous contents."""
    jpegs = synth.synth_jpeg_set(4)
    full_y, full_c, full_q, full_f = dm.decode_batch(jpegs, 64, 64, nthreads=2)
    last = [10, 63, 0, 31]
    y = torch.full((4, 64, 64, 64), -7, dtype=torch.int16)
    c = torch.full((4, 2, 32, 32, 64), -7, dtype=torch.int16)
    q = torch.zeros((4, 3, 64), dtype=torch.int16)
    y2, c2, q2, f2 = dm.decode_batch(jpegs, 64, 64, nthreads=2, out=(y, c, q), last_rows=last)
    assert torch.equal(q2, full_q)
    for i, r in enumerate(last):
        mcu = r // 2                                   # 4:2:0: one MCU row = two luma block rows, one chroma block row
        assert torch.equal(y2[i, :2 * mcu + 2], full_y[i, :2 * mcu + 2])
        assert torch.equal(c2[i, :, :mcu + 1], full_c[i, :, :mcu + 1])
        assert bool((y2[i, 2 * mcu + 2:] == -7).all()) and bool((c2[i, :, mcu + 1:] == -7).all())
    assert int(f2[1]) == int(full_f[1])               # whole image decoded: same clamp flag
    with pytest.raises(ValueError):
        dm.decode_batch(jpegs, 64, 64, last_rows=[1, 2])


def test_decode_coeff_and_quantize_at_quality_round_trip_through_libjpeg():
    """B1 completeness (dct_manip.cpp:315-375, 485-576): quantize_at_quality = libjpeg's compressor + the coefficient reader;
    decode_coeff = the coefficient writer + libjpeg's decompressor.  Pixels -> coefficients -> pixels must equal what libjpeg itself
    decodes from the file it wrote, and the tables must be libjpeg's tables for that quality."""
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:96, 0:128]
    img = np.stack([127 + 80 * np.sin(xx / 11.0 + c) * np.cos(yy / 7.0) + rng.normal(0, 10, (96, 128)) for c in range(3)], 0)
    pix = torch.from_numpy(np.clip(img, 0, 255).astype(np.uint8))
    for quality in (100, 60):
        dims, quant, y, cbcr = dm.quantize_at_quality(pix, quality)
        assert tuple(y.shape) == (1, 12, 16, 8, 8) and tuple(cbcr.shape) == (2, 6, 8, 8, 8) and dims[0].tolist() == [96, 128]
        assert torch.equal(quant[:2], dm.quality_tables(quality)) and torch.equal(quant[1], quant[2])
        rgb = dm.decode_coeff(dims, quant, y, cbcr)
        buf = io.BytesIO()
        Image.fromarray(pix.permute(1, 2, 0).numpy()).save(buf, "JPEG", quality=quality, subsampling=2)
        ref = torch.from_numpy(np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).copy()).permute(2, 0, 1)
        assert rgb.dtype == torch.uint8 and torch.equal(rgb, ref)
        assert torch.equal(dm.decode_coeff(dims, quant * 0 + 1, y, cbcr, quality=quality), ref)      # tables from `quality`
        assert float((rgb.float() - pix.float()).abs().mean()) < 12.0       # lossy (per-channel noise vs 4:2:0), but the same picture
    g = dm.quantize_at_quality(pix[:1], 90)
    assert g[3] is None and tuple(dm.decode_coeff(g[0], g[1], g[2]).shape) == (1, 96, 128)
