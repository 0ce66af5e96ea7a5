"""Device-side UFI construction (urmb_build.cu) and UFI file writing (format of ufindexio.cpp:15-49)."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import engine


def get_prime(n: int) -> int:
    """prime.cpp:11 over primes.h: first entry >= n of the table "first prime >= x", x = 100, x <- x*100//95."""
    def is_prime(v):
        if v < 2:
            return False
        if v % 2 == 0:
            return v == 2
        d, s = v - 1, 0
        while d % 2 == 0:
            d //= 2
            s += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if a % v == 0:
                continue
            x = pow(a, d, v)
            if x in (1, v - 1):
                continue
            for _ in range(s - 1):
                x = x * x % v
                if x == v - 1:
                    break
            else:
                return False
        return True

    x = 100
    for _ in range(410):
        p = x
        while not is_prime(p):
            p += 1
        if p >= n:
            return p
        x = x * 100 // 95
    raise ValueError("GetPrime overflow")


def fasta_bytes(names, lens, cols=60):
    """Size of the FASTA file the genome would occupy (cmd_make_ufi sizes the table from it, ufindexio.cpp:138)."""
    return sum(len(nm) + 2 + int(L) + (int(L) + cols - 1) // cols for nm, L in zip(names, lens))


def slot_count_for(names, lens, load_factor=0.6):
    return get_prime(int(fasta_bytes(names, lens) / load_factor))


def build_index_device(d_seq_ptr: int, seq_data_size: int, slot_count: int, d_blob_ptr: int, word_length=24, max_ix=32):
    """Runs the builder kernels on the current CUDA device. Returns dict(indexed, truncated, seconds); the blob is
    byte-identical to the reference's when `truncated` (segments that need long links / truncated lists) is 0."""
    L = engine.lib()
    L.urmb_build_index_device.restype = C.c_int
    L.urmb_build_index_device.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                          C.c_void_p]
    L.urmb_build_last_error.restype = C.c_char_p
    stats = np.zeros(3, dtype=np.uint64)
    rc = L.urmb_build_index_device(d_seq_ptr, seq_data_size, slot_count, word_length, max_ix, d_blob_ptr,
                                   stats.ctypes.data)
    if rc != 0:
        raise engine.UrmbError(rc, L.urmb_build_last_error().decode())
    return {"indexed": int(stats[0]), "truncated": int(stats[1]), "seconds": int(stats[2]) / 1e6}


def ufi_header(names, lens, offsets, seq_data_size, slot_count, word_length=24, max_ix=32) -> bytes:
    h = struct.pack("<IIIIQI", 0x55464931, word_length, max_ix, seq_data_size, slot_count, len(names))
    for nm, L, off in zip(names, lens, offsets):
        b = nm.encode()
        h += struct.pack("<III", int(L), int(off), len(b)) + b
    return h + struct.pack("<I", 0x55464932)


def write_ufi(path, names, lens, offsets, seq_data_size, slot_count, blob_chunks, seq_chunks, word_length=24,
              max_ix=32):
    """blob_chunks / seq_chunks: iterables of bytes-like objects (so a 27 GB blob can stream from the device)."""
    with open(path, "wb") as f:
        f.write(ufi_header(names, lens, offsets, seq_data_size, slot_count, word_length, max_ix))
        for c in blob_chunks:
            f.write(c)
        f.write(struct.pack("<I", 0x55464933))
        for c in seq_chunks:
            f.write(c)
        f.write(struct.pack("<I", 0x55464935))
