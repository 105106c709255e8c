"""bonej2_b200.rawio against the behaviour DatasetUtilTest.java pins for DatasetUtil.loadRaw3D / readRawBytes /
saveAsRaw (Modern/utilities/src/test/java/org/bonej/utilities/DatasetUtilTest.java:96-215, 301-430): argument
checks, header skipping, truncated files, 0/1 <-> 0/255, byte order round trips -- plus the file-to-file
Thickness driver on the GPU."""
import numpy as np
import pytest

from bonej2_b200 import rawio
from tests import shapes


def _write(tmp_path, name, data: bytes):
    p = tmp_path / name
    p.write_bytes(data)
    return str(p)


@pytest.mark.parametrize("dims", [(0, 10, 10), (10, 0, 10), (10, 10, 0), (-1, 2, 2)])
def test_load_invalid_dimensions(tmp_path, dims):                     # :96-118
    p = _write(tmp_path, "a.raw", bytes(1000))
    with pytest.raises(ValueError, match="Dimensions must be positive integers."):
        rawio.load_raw3d(p, 0, *dims, "uint8")


def test_load_missing_or_null_file(tmp_path):                         # :120-132
    with pytest.raises(ValueError, match="File does not exist"):
        rawio.load_raw3d(str(tmp_path / "nope.raw"), 0, 2, 2, 2, "uint8")
    with pytest.raises(ValueError, match="File does not exist"):
        rawio.load_raw3d(None, 0, 2, 2, 2, "uint8")


def test_load_unsupported_pixel_type(tmp_path):                       # :134-141
    p = _write(tmp_path, "a.raw", bytes(64))
    with pytest.raises(ValueError, match="Unsupported pixel type: int64"):
        rawio.load_raw3d(p, 0, 2, 2, 2, "int64")


def test_load_file_too_small_and_offset_beyond(tmp_path):             # :143-159
    p = _write(tmp_path, "a.raw", bytes(10))
    with pytest.raises(IOError, match=r"File too small: 10 bytes available, but 27 needed \(offset=0 \+ data=27\)"):
        rawio.load_raw3d(p, 0, 3, 3, 3, "uint8")
    with pytest.raises(IOError, match="File too small"):
        rawio.load_raw3d(p, 100, 1, 1, 1, "uint8")


def test_read_raw_bytes(tmp_path):                                    # :161-215
    p = _write(tmp_path, "r.raw", bytes([1, 2, 3, 4, 5]))
    assert rawio.read_raw_bytes(p, 0, 5).tolist() == [1, 2, 3, 4, 5]
    header, payload = bytes([9] * 7), bytes(range(10, 30))
    p = _write(tmp_path, "h.raw", header + payload)
    assert rawio.read_raw_bytes(p, 7, 20).tobytes() == payload
    big = np.random.default_rng(0).integers(0, 256, 1 << 20, dtype=np.uint8).tobytes()
    p = _write(tmp_path, "big.raw", big)
    assert rawio.read_raw_bytes(p, 0, len(big)).tobytes() == big
    assert rawio.read_raw_bytes(p, 0, 0).size == 0
    with pytest.raises(IOError, match="Unexpected EOF: read 5 of 8 expected pixel bytes."):
        rawio.read_raw_bytes(_write(tmp_path, "t.raw", bytes(5)), 0, 8)
    with pytest.raises(IOError, match="Unexpected end of stream while skipping header."):
        rawio.read_raw_bytes(_write(tmp_path, "s.raw", bytes(3)), 8, 1)


def test_blocked_reads_cross_block_boundaries(tmp_path, monkeypatch):
    monkeypatch.setattr(rawio, "BLOCK_BYTES", 1000)
    data = np.random.default_rng(1).integers(0, 2, 7 * 11 * 13, dtype=np.uint8)
    p = _write(tmp_path, "b.raw", bytes(5) + data.tobytes())
    v = rawio.load_raw3d(p, 5, 13, 11, 7, "uint8", zero_one_binary=True, spacing_x=0.1, spacing_y=0.2, spacing_z=0.3)
    assert v.data.shape == (7, 11, 13) and v.name == "b.raw" and v.spacing == (0.1, 0.2, 0.3)
    assert np.array_equal(v.data.reshape(-1), data * 255)


def test_zero_one_rejects_other_values(tmp_path):                     # DatasetUtil.java:136-147
    p = _write(tmp_path, "z.raw", bytes([0, 1, 1, 0, 255, 1, 0, 1]))
    with pytest.raises(ValueError, match=r"Non-binary value detected at index 4: -1 \(0xFF\)"):
        rawio.load_raw3d(p, 0, 2, 2, 2, "uint8", zero_one_binary=True)
    v = rawio.load_raw3d(p, 0, 2, 2, 2, "uint8", zero_one_binary=False)           # untouched without the flag
    assert v.data.reshape(-1).tolist() == [0, 1, 1, 0, 255, 1, 0, 1]
    v16 = rawio.load_raw3d(p, 0, 2, 2, 1, "uint16", zero_one_binary=True)         # ignored for 16-bit
    assert v16.data.dtype == np.uint16


@pytest.mark.parametrize("ptype,dtype", [("uint16", np.uint16), ("int16", np.int16), ("float32", np.float32)])
@pytest.mark.parametrize("order,le", [("Little-endian", True), ("Big-endian", False)])
def test_round_trip_multibyte(tmp_path, ptype, dtype, order, le):     # :325-430
    w, h, d = 5, 4, 3
    vals = (np.arange(w * h * d) * 37 - 500).astype(dtype).reshape(d, h, w)
    p = str(tmp_path / "m.raw")
    rawio.save_as_raw(vals, p, little_endian=le)
    raw = np.fromfile(p, dtype=np.dtype(dtype).newbyteorder("<" if le else ">"))
    assert np.array_equal(raw.astype(dtype).reshape(d, h, w), vals)
    back = rawio.load_raw3d(p, 0, w, h, d, ptype, byte_order=order)
    assert back.data.dtype == dtype and np.array_equal(back.data, vals)


def test_save_uint8_and_zero_one(tmp_path):                           # :301-323, DatasetUtil.java:404-414
    vol = shapes.random_blobs((4, 5, 6), 1)
    p = str(tmp_path / "u.raw")
    rawio.save_as_raw(vol, p)
    assert np.array_equal(np.fromfile(p, np.uint8).reshape(vol.shape), vol)
    rawio.save_as_raw(vol, p, zero_one_binary=True)
    assert np.array_equal(np.fromfile(p, np.uint8).reshape(vol.shape), vol // 255)
    assert np.array_equal(rawio.load_raw3d(p, 0, 6, 5, 4, "uint8", zero_one_binary=True).data, vol)
    bad = vol.copy(); bad[0, 0, 0] = 7
    with pytest.raises(ValueError, match="Invalid binary value: 7"):
        rawio.save_as_raw(bad, p, zero_one_binary=True)
    with pytest.raises(ValueError, match="Expected 2D or 3D Dataset"):
        rawio.save_as_raw(np.zeros((2, 2, 2, 2), np.uint8), p)
    with pytest.raises(NotImplementedError, match="Unsupported pixel type"):
        rawio.save_as_raw(np.zeros((2, 2), np.float64), p)


@pytest.mark.gpu
@pytest.mark.parametrize("choice", ["Both", "Trabecular separation"])
def test_thickness_raw_file_to_file(tmp_path, oracle, choice):
    vol = shapes.random_blobs((30, 26, 34), 8)
    src = str(tmp_path / "vol01.raw")
    with open(src, "wb") as f:
        f.write(b"HDR!" * 4)
        f.write((vol // 255).tobytes())                               # stored as 0 / 1 behind a 16-byte header
    d, h, w = vol.shape
    res = rawio.thickness_raw(src, w, h, d, out_prefix=str(tmp_path / "map"), header_offset=16, zero_one_binary=True,
                              map_choice=choice, mask=True, spacing=0.25)
    keys = ["Tb.Th", "Tb.Sp"] if choice == "Both" else ["Tb.Sp"]
    assert sorted(res) == sorted(keys)
    for key in keys:
        ref = oracle.process_image(vol, inverse=(key == "Tb.Sp"), mask=True, pixel_width=0.25)
        got = rawio.load_raw3d(str(tmp_path / f"map_{key}.raw"), 0, w, h, d, "float32").data
        assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
        ok = ~np.isnan(got)
        assert np.allclose(got[ok], ref["map"][ok], rtol=1e-5, atol=0)
        st = res[key][0]
        assert st.count == ref["stats"].count
        assert abs(st.mean - ref["stats"].mean) <= 1e-5 * abs(ref["stats"].mean)


@pytest.mark.gpu
def test_thickness_raw_refuses_non_binary(tmp_path):
    p = str(tmp_path / "g.raw")
    np.arange(27, dtype=np.uint8).tofile(p)
    with pytest.raises(Exception, match="binary"):
        rawio.thickness_raw(p, 3, 3, 3)
