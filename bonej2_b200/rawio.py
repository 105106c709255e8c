"""Raw volume I/O compatible with org.bonej.utilities.DatasetUtil, and a Thickness
driver that streams a raw file through the CUDA path.

`load_raw3d` / `save_as_raw` follow `DatasetUtil.loadRaw3D` (Modern/utilities/.../DatasetUtil.java:99-177) and
`DatasetUtil.saveAsRaw` (:380-478): same parameters, same checks and messages, same byte layout ([z][y][x], x
fastest, optional header, 0/1 <-> 0/255 for 8-bit binary data, little- or big-endian 16-bit and float
samples).  The reference reads the file into one Java `byte[]`, which caps it at 2 GiB; here the file is read
in plane blocks, so a 2048^3 phantom (8 GiB) loads too.

`thickness_raw` is the file-to-file form of the Thickness command: the u8 volume is read block by block
straight into page-locked memory (bjt_host_alloc), the 0/1 -> 0/255 mapping is applied on the way, the
C-ABI call (bjt_thickness_both_u8 or bjt_thickness_u8) runs on the GPU, and the float32 maps are written
from page-locked memory.  No CPU path: without the CUDA library the call fails.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

PIXEL_TYPES = {"uint8": (np.uint8, 1), "uint16": (np.uint16, 2), "int16": (np.int16, 2), "float32": (np.float32, 4)}
BLOCK_BYTES = 64 << 20                     # file reads / writes are issued in pieces of about this size


@dataclass
class RawVolume:
    data: np.ndarray                        # [depth][height][width]
    name: str
    spacing: tuple                          # (x, y, z) in mm, as DatasetUtil sets on the axes
    unit: str = "mm"


def _dtype(pixel_type: str, byte_order: str):
    pt = pixel_type.lower()
    if pt not in PIXEL_TYPES:
        raise ValueError(f"Unsupported pixel type: {pixel_type}")
    dt, bpp = PIXEL_TYPES[pt]
    order = ">" if str(byte_order).lower() == "big-endian" else "<"       # anything else is little-endian (:150-152)
    return np.dtype(dt).newbyteorder(order) if bpp > 1 else np.dtype(dt), bpp


def _check_file(path, header_offset, width, height, depth, bpp):
    if width <= 0 or height <= 0 or depth <= 0:
        raise ValueError("Dimensions must be positive integers.")
    if path is None or not os.path.exists(path):
        raise ValueError(f"File does not exist: {path}")
    data_bytes = width * height * depth * bpp
    need = header_offset + data_bytes
    actual = os.path.getsize(path)
    if actual < need:
        raise IOError(f"File too small: {actual} bytes available, but {need} needed (offset={header_offset} + data={data_bytes}).")
    return data_bytes


def read_raw_bytes(path, header_offset: int, data_length: int, out=None) -> np.ndarray:
    """DatasetUtil.readRawBytes (:188-212): `data_length` bytes after the header, in blocks; `out` may be any
    writable u8 buffer of that length (e.g. page-locked memory)."""
    buf = np.empty(data_length, np.uint8) if out is None else out
    with open(path, "rb", buffering=0) as f:
        f.seek(0, os.SEEK_END)
        if f.tell() < header_offset:
            raise IOError("Unexpected end of stream while skipping header.")
        f.seek(header_offset)
        off = 0
        mv = memoryview(buf).cast("B")
        while off < data_length:
            n = f.readinto(mv[off:min(data_length, off + BLOCK_BYTES)])
            if not n:
                raise IOError(f"Unexpected EOF: read {off} of {data_length} expected pixel bytes.")
            off += n
    return buf


def zero_one_to_255(buf: np.ndarray):
    """in place 0 -> 0, 1 -> 255; anything else is an error naming the first offender (:136-147)"""
    step = BLOCK_BYTES
    flat = buf.reshape(-1)
    for a in range(0, flat.size, step):
        blk = flat[a:a + step]
        bad = np.flatnonzero(blk & 0xFE)
        if bad.size:
            i = a + int(bad[0])
            v = int(flat[i])
            sv = v - 256 if v > 127 else v
            raise ValueError(f"Non-binary value detected at index {i}: {sv} (0x{v:02X})")
        np.multiply(blk, 255, out=blk)


def load_raw3d(path, header_offset, width, height, depth, pixel_type, zero_one_binary=False, byte_order="Little-endian",
               spacing_x=1.0, spacing_y=1.0, spacing_z=1.0) -> RawVolume:
    dt, bpp = _dtype(pixel_type, byte_order)
    nbytes = _check_file(path, header_offset, width, height, depth, bpp)
    raw = read_raw_bytes(path, header_offset, nbytes)
    if zero_one_binary and bpp == 1:
        zero_one_to_255(raw)
    data = raw.view(dt).reshape(depth, height, width)
    if bpp > 1:
        data = data.astype(dt.newbyteorder("="))
    return RawVolume(data, os.path.basename(str(path)), (spacing_x, spacing_y, spacing_z))


def save_as_raw(data: np.ndarray, path, little_endian=True, zero_one_binary=False):
    """DatasetUtil.saveAsRaw: a 2-D or 3-D array of u8 / i8 / bool, u16 / i16 or f32, samples in file order.
    zero_one_binary writes 8-bit 0 / 255 data as 0 / 1 and refuses any other value."""
    a = np.asarray(data)
    if a.ndim < 2 or a.ndim > 3:
        raise ValueError("Expected 2D or 3D Dataset")
    if a.dtype == np.bool_:
        a = a.astype(np.uint8) * (1 if zero_one_binary else 255)
        zero_one_binary = False
    if a.dtype in (np.uint8, np.int8):
        out_dt = a.dtype
    elif a.dtype in (np.uint16, np.int16, np.float32):
        out_dt = a.dtype.newbyteorder("<" if little_endian else ">")
    else:
        raise NotImplementedError(f"Unsupported pixel type: {a.dtype}")
    flat = np.ascontiguousarray(a).reshape(-1)
    step = max(1, BLOCK_BYTES // flat.dtype.itemsize)
    with open(path, "wb") as f:
        for s in range(0, flat.size, step):
            blk = flat[s:s + step]
            if zero_one_binary and blk.dtype.itemsize == 1:
                u = blk.view(np.uint8)
                bad = np.flatnonzero((u != 0) & (u != 255))
                if bad.size:
                    raise ValueError(f"Invalid binary value: {int(u[bad[0]])}")
                blk = u & 1
            f.write(blk.astype(out_dt, copy=False).tobytes())


# ----------------------------------------------------------------------------------------------------------

def thickness_raw(path, width, height, depth, out_prefix=None, header_offset=0, zero_one_binary=False,
                  map_choice="Both", mask=True, spacing=1.0, device=0, little_endian=True):
    """Thickness on a raw 8-bit binary volume, file to file.

    Returns {"Tb.Th": (stats, timing), "Tb.Sp": (...)} for the chosen maps; when `out_prefix` is given the
    float32 maps go to `<out_prefix>_Tb.Th.raw` / `_Tb.Sp.raw` in DatasetUtil's layout (NaN = no value).
    Anything but 0 / 255 after the optional 0/1 mapping is refused like the command does ("Need a binary image")."""
    from . import _native
    if map_choice not in ("Trabecular thickness", "Trabecular separation", "Both"):
        raise ValueError("Unexpected map choice")
    nbytes = _check_file(path, header_offset, width, height, depth, 1)
    L = _native.lib()
    ctx = _native.Context(device)
    N = nbytes
    h_in = L.bjt_host_alloc(N)
    h_a = L.bjt_host_alloc(4 * N)
    h_b = L.bjt_host_alloc(4 * N) if map_choice == "Both" else None
    if not h_in or not h_a or (map_choice == "Both" and not h_b):
        raise MemoryError("could not allocate page-locked host buffers")
    try:
        vin = np.ctypeslib.as_array(ctypes.cast(h_in, ctypes.POINTER(ctypes.c_uint8)), shape=(N,))
        read_raw_bytes(path, header_offset, N, out=vin)
        if zero_one_binary:
            zero_one_to_255(vin)
        res = {}
        if map_choice == "Both":
            (s0, t0), (s1, t1) = ctx.thickness_both_raw(h_in, width, height, depth, mask, 128, spacing, h_a, h_b)
            res = {"Tb.Th": (s0, t0, h_a), "Tb.Sp": (s1, t1, h_b)}
        else:
            inverse = map_choice == "Trabecular separation"
            vol = vin.reshape(depth, height, width)
            out = np.ctypeslib.as_array(ctypes.cast(h_a, ctypes.POINTER(ctypes.c_float)), shape=(depth, height, width))
            _, st, tm = ctx.thickness(vol, inverse=inverse, mask=mask, pixel_width=spacing, out=out)
            res = {"Tb.Sp" if inverse else "Tb.Th": (st, tm, h_a)}
        ret = {}
        for key, (st, tm, ptr) in res.items():
            if out_prefix is not None:
                m = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(depth, height, width))
                save_as_raw(m, f"{out_prefix}_{key}.raw", little_endian=little_endian)
            ret[key] = (st, tm)
        return ret
    finally:
        L.bjt_host_free(h_in)
        L.bjt_host_free(h_a)
        if h_b:
            L.bjt_host_free(h_b)
        ctx.close()
