"""Minimal FITS image writer / reader for the state dumps of GridSolver:save (hydro/solver/gridsolver.lua:1410-1470).

The reference saves each buffer through its `image` library's FITS back end as a 3-axis image of `real`s:
NAXIS1 = width * depth, NAXIS2 = height, NAXIS3 = channels (planar, one channel per state variable), where element
(ch, i, j, k) sits at  i + width * (k + depth * (j + height * ch))  -- gridsolver.lua:1426-1429, "interleave the depth and
the width".  FITS itself (NASA/IAU standard 4.0): 2880-byte blocks, 80-character header cards, big-endian data,
BITPIX = -64 for double and -32 for float.  `image` is an un-vendored dependency of the reference, so the exact set of
optional cards it writes is not pinned; a reader only needs the mandatory ones, which is what is written and required here.
"""
import numpy as np

BLOCK = 2880


def _card(key, value=None, comment=""):
    if value is None:
        s = key
    elif isinstance(value, bool):
        s = "%-8s= %20s" % (key, "T" if value else "F")
    elif isinstance(value, int):
        s = "%-8s= %20d" % (key, value)
    else:
        s = "%-8s= %20s" % (key, value)
    if comment:
        s += " / " + comment
    return s[:80].ljust(80)


def write_image(filename, data):
    """data: array [NAXIS3][NAXIS2][NAXIS1] (C order) of float64 or float32."""
    data = np.asarray(data)
    if data.dtype not in (np.float64, np.float32):
        raise TypeError("FITS state dumps hold float64 or float32")
    if data.ndim != 3:
        raise ValueError("expected [channels][height][width*depth]")
    bitpix = -64 if data.dtype == np.float64 else -32
    cards = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", bitpix, "number of bits per data pixel"),
             _card("NAXIS", 3, "number of data axes"), _card("NAXIS1", int(data.shape[2])), _card("NAXIS2", int(data.shape[1])),
             _card("NAXIS3", int(data.shape[0])), _card("EXTEND", True), _card("END")]
    header = "".join(cards)
    header += " " * (-len(header) % BLOCK)
    raw = data.astype(data.dtype.newbyteorder(">"), copy=False).tobytes()
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(raw)
        f.write(b"\0" * (-len(raw) % BLOCK))


def read_image(filename):
    """-> array [NAXIS3][NAXIS2][NAXIS1] in native byte order (float64 / float32); accepts any conforming primary image HDU."""
    with open(filename, "rb") as f:
        cards = {}
        done = False
        while not done:
            block = f.read(BLOCK)
            if len(block) < BLOCK:
                raise ValueError("%s: truncated FITS header" % filename)
            for n in range(0, BLOCK, 80):
                c = block[n:n + 80].decode("ascii")
                key = c[:8].strip()
                if key == "END":
                    done = True
                    break
                if c[8:10] == "= ":
                    cards[key] = c[10:].split("/")[0].strip()
        if cards.get("SIMPLE") != "T":
            raise ValueError("%s: not a FITS file" % filename)
        bitpix = int(cards["BITPIX"])
        if bitpix not in (-64, -32):
            raise ValueError("%s: BITPIX %d is not a real-valued state dump" % (filename, bitpix))
        naxis = int(cards["NAXIS"])
        shape = [int(cards["NAXIS%d" % (n + 1)]) for n in range(naxis)][::-1]
        while len(shape) < 3:
            shape.insert(0, 1)
        dt = np.dtype(">f8" if bitpix == -64 else ">f4")
        count = int(np.prod(shape))
        raw = f.read(count * dt.itemsize)
        if len(raw) < count * dt.itemsize:
            raise ValueError("%s: truncated FITS data" % filename)
        scale, zero = float(cards.get("BSCALE", 1.)), float(cards.get("BZERO", 0.))
        a = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
        if scale != 1. or zero != 0.:
            a = a * scale + zero
        return a.reshape(shape)


def state_to_image(U, real=np.float64):
    """AoS state [Sz][Sy][Sx][channels] -> the reference's planar image [channels][Sy][Sz * Sx] (gridsolver.lua:1426-1445)."""
    U = np.asarray(U)
    d, h, w, c = U.shape
    return np.ascontiguousarray(U.transpose(3, 1, 0, 2).reshape(c, h, d * w).astype(real))


def image_to_state(img, gridSize):
    """inverse of state_to_image for gridSize = (Sx, Sy, Sz)."""
    w, h, d = gridSize
    c = img.shape[0]
    if img.shape[1] != h or img.shape[2] != w * d:
        raise ValueError("FITS image %s does not match the grid %s" % (img.shape, (gridSize,)))
    return np.ascontiguousarray(img.reshape(c, h, d, w).transpose(2, 1, 3, 0).astype(np.float64))
