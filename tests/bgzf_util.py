"""BGZF writer for tests and benches (SAM spec 4.1): independent gzip members of <= 0xff00 input bytes, each with a
'BC' extra subfield, followed by the 28-byte EOF marker.  `level` / `strategy` choose what kind of DEFLATE blocks zlib
emits (level 0: stored; Z_FIXED: fixed Huffman; default: dynamic)."""
import struct
import zlib

EOF_MARKER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_member(chunk: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    payload = co.compress(chunk) + co.flush()
    bsize = len(payload) + 25  # total member size - 1
    assert bsize < 65536
    hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
    return hdr + payload + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))


def bgzf_compress(data: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY, block: int = 0xFF00, eof: bool = True) -> bytes:
    data = bytes(data)
    if level == 0:
        block = min(block, 0xFF00 - 64)
    out = [bgzf_member(data[i:i + block], level, strategy) for i in range(0, len(data), block)]
    if eof:
        out.append(EOF_MARKER)
    return b"".join(out)
