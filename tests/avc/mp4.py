"""Minimal MP4 (ISO BMFF) MUXER for the streams tests/avc/stream.py writes: one video track, avc1 sample entry with an
avcC box, one sample (= one IDR access unit, NAL units length-prefixed) per picture. TEST TOOLING ONLY: it exists so that
the container layer of the CPU host (dryv_mp4_scan / dryv_mp4_parse, the counterpart of the reference's src/video/atom/**
and src/video/sample/**) can be fed the kind of file `dryv <file>` opens, and so that libavformat can confirm the file is
a valid MP4. Box layouts per ISO/IEC 14496-12 and 14496-15 (avcC)."""
from __future__ import annotations

import struct


def box(kind: bytes, payload: bytes) -> bytes:
    return struct.pack(">I", 8 + len(payload)) + kind + payload


def full(kind: bytes, version: int, flags: int, payload: bytes) -> bytes:
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


def split_annexb(data: bytes):
    """-> list of NAL units (without start codes)."""
    out, i, n = [], 0, len(data)
    starts = []
    while i + 3 <= n:
        if data[i] == 0 and data[i + 1] == 0 and data[i + 2] == 1:
            starts.append(i + 3)
            i += 3
        else:
            i += 1
    for k, s in enumerate(starts):
        e = starts[k + 1] - 3 if k + 1 < len(starts) else n
        while e > s and data[e - 1] == 0:
            e -= 1
        out.append(data[s:e])
    return out


def mux(annexb: bytes, width: int, height: int, timescale: int = 25) -> bytes:
    nals = split_annexb(annexb)
    sps = [u for u in nals if u[0] & 31 == 7]
    pps = [u for u in nals if u[0] & 31 == 8]
    samples = [struct.pack(">I", len(u)) + u for u in nals if u[0] & 31 == 5]   # one IDR slice NAL per sample
    assert sps and pps and samples
    avcc = bytes([1, sps[0][1], sps[0][2], sps[0][3], 0xFC | 3, 0xE0 | 1]) + struct.pack(">H", len(sps[0])) + sps[0] + \
        bytes([1]) + struct.pack(">H", len(pps[0])) + pps[0]
    avc1 = struct.pack(">6sH", b"\0" * 6, 1) + struct.pack(">HHIII", 0, 0, 0, 0, 0) + struct.pack(">HH", width, height) + \
        struct.pack(">IIIH", 0x00480000, 0x00480000, 0, 1) + b"\0" * 32 + struct.pack(">Hh", 0x18, -1) + box(b"avcC", avcc)
    stsd = full(b"stsd", 0, 0, struct.pack(">I", 1) + box(b"avc1", avc1))
    n = len(samples)
    stts = full(b"stts", 0, 0, struct.pack(">III", 1, n, 1))
    stss = full(b"stss", 0, 0, struct.pack(">I", n) + b"".join(struct.pack(">I", i + 1) for i in range(n)))
    stsc = full(b"stsc", 0, 0, struct.pack(">IIII", 1, 1, n, 1))       # one chunk holding every sample
    stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, n) + b"".join(struct.pack(">I", len(s)) for s in samples))
    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 512) + b"isomiso2avc1mp41")

    def moov(mdat_payload_offset: int) -> bytes:
        stco = full(b"stco", 0, 0, struct.pack(">II", 1, mdat_payload_offset))
        stbl = box(b"stbl", stsd + stts + stss + stsc + stsz + stco)
        dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
        minf = box(b"minf", full(b"vmhd", 0, 1, struct.pack(">HHHH", 0, 0, 0, 0)) + dinf + stbl)
        mdhd = full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, n, 0x55C4, 0))
        hdlr = full(b"hdlr", 0, 0, struct.pack(">I4sIII", 0, b"vide", 0, 0, 0) + b"VideoHandler\0")
        mdia = box(b"mdia", mdhd + hdlr + minf)
        matrix = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)
        tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, n) + struct.pack(">IIhhhH", 0, 0, 0, 0, 0, 0) + matrix +
                    struct.pack(">II", width << 16, height << 16))
        trak = box(b"trak", tkhd + mdia)
        mvhd = full(b"mvhd", 0, 0, struct.pack(">IIIIIH", 0, 0, timescale, n, 0x10000, 0x100) + b"\0" * 10 + matrix +
                    b"\0" * 24 + struct.pack(">I", 2))
        return box(b"moov", mvhd + trak)

    probe = moov(0)
    offset = len(ftyp) + len(probe) + 8
    return ftyp + moov(offset) + box(b"mdat", b"".join(samples))
