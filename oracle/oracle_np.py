"""Independent numpy restatement of the reference algorithm (TEST INFRASTRUCTURE ONLY).

Second opinion for oracle/fftconv_oracle.c: same state machines, but the FFT is numpy's
pocketfft in single precision (np.fft.rfft / irfft on float32 keep complex64 / float32 in
numpy >= 2.0) instead of the C oracle's radix-2.  Two unrelated FFTs agreeing to ~1e-6·RMS is
what stands in for the unavailable realfft/rustfft bit patterns.

Citations are to /root/reference/src/fft_convolver.rs unless stated.
"""
from __future__ import annotations

import math

import numpy as np


class Panic(RuntimeError):
    pass


def next_power_of_two(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


def rfft_f32(x: np.ndarray) -> np.ndarray:
    out = np.fft.rfft(x.astype(np.float32, copy=False))
    assert out.dtype == np.complex64
    return out


def irfft_f32(X: np.ndarray, n: int) -> np.ndarray:
    """Fft::inverse (:41-49): unnormalised C2R then `/= len as f32`.  numpy's irfft already
    normalises by 1/n; for power-of-two n that is the same scaling up to the last bit."""
    out = np.fft.irfft(X.astype(np.complex64, copy=False), n=n)
    assert out.dtype == np.float32
    return out


def cmac(result: np.ndarray, a: np.ndarray, b: np.ndarray) -> None:
    """complex_multiply_accumulate (:62-74), each operation rounded to f32 separately."""
    ar, ai, br, bi = a.real, a.imag, b.real, b.imag
    pr = (ar * br).astype(np.float32) - (ai * bi).astype(np.float32)
    pi = (ar * bi).astype(np.float32) + (ai * br).astype(np.float32)
    result.real += pr.astype(np.float32)
    result.imag += pi.astype(np.float32)


class FFTConvolverNP:
    """:86-307"""

    def __init__(self):
        self.ir_len = self.block_size = self.seg_count = self.active_seg_count = 0
        self.current = self.input_buffer_fill = 0
        self.segments = self.segments_ir = None

    @classmethod
    def init(cls, ir, block_size, max_response_length):
        ir = np.asarray(ir, dtype=np.float32)
        if max_response_length < ir.size:
            raise Panic("max_response_length must be at least the length of the initial impulse response")
        s = cls()
        padded = np.zeros(max_response_length, np.float32)
        padded[: ir.size] = ir
        s.ir_len = padded.size
        B = s.block_size = next_power_of_two(block_size)
        s.seg_count = math.ceil(s.ir_len / B)
        s.active_seg_count = s.seg_count
        K = B + 1
        s.segments = np.zeros((s.seg_count, K), np.complex64)
        s.segments_ir = np.zeros((s.seg_count, K), np.complex64)
        for i in range(s.seg_count):
            buf = np.zeros(2 * B, np.float32)
            chunk = padded[i * B : (i + 1) * B]
            buf[: chunk.size] = chunk
            s.segments_ir[i] = rfft_f32(buf)
        s.pre_multiplied = np.zeros(K, np.complex64)
        s.conv = np.zeros(K, np.complex64)
        s.overlap = np.zeros(B, np.float32)
        s.input_buffer = np.zeros(B, np.float32)
        s.fft_buffer = np.zeros(2 * B, np.float32)
        return s

    def update(self, response):
        response = np.asarray(response, dtype=np.float32)
        if response.size > self.ir_len:
            raise Panic("New impulse response is longer than initialized length")
        if self.ir_len == 0:
            return
        B = self.block_size
        self.fft_buffer[:] = 0
        self.conv[:] = 0
        self.pre_multiplied[:] = 0
        self.overlap[:] = 0
        self.active_seg_count = math.ceil(response.size / B)
        for i in range(self.active_seg_count):
            buf = np.zeros(2 * B, np.float32)
            chunk = response[i * B : (i + 1) * B]
            buf[: chunk.size] = chunk
            self.segments_ir[i] = rfft_f32(buf)
        self.segments_ir[self.active_seg_count :] = 0

    def process(self, inp, out):
        if self.active_seg_count == 0:
            out[:] = 0
            return
        if len(inp) < len(out):
            raise Panic("range end index out of range")
        B = self.block_size
        processed = 0
        while processed < len(out):
            was_empty = self.input_buffer_fill == 0
            n = min(len(out) - processed, B - self.input_buffer_fill)
            pos = self.input_buffer_fill
            self.input_buffer[pos : pos + n] = inp[processed : processed + n]
            self.fft_buffer[:B] = self.input_buffer
            self.fft_buffer[B:] = 0
            self.segments[self.current] = rfft_f32(self.fft_buffer)
            if was_empty:
                self.pre_multiplied[:] = 0
                for i in range(1, self.active_seg_count):
                    cmac(self.pre_multiplied, self.segments_ir[i],
                         self.segments[(self.current + i) % self.active_seg_count])
            self.conv[:] = self.pre_multiplied
            cmac(self.conv, self.segments[self.current], self.segments_ir[0])
            c = self.conv.copy()
            c.imag[0] = 0
            c.imag[-1] = 0
            self.fft_buffer[:] = irfft_f32(c, 2 * B)
            out[processed : processed + n] = self.fft_buffer[pos : pos + n] + self.overlap[pos : pos + n]
            self.input_buffer_fill += n
            if self.input_buffer_fill == B:
                self.input_buffer[:] = 0
                self.input_buffer_fill = 0
                self.overlap[:] = self.fft_buffer[B:]
                self.current = self.current - 1 if self.current > 0 else self.active_seg_count - 1
            processed += n

    def reset(self):
        if self.segments is None:
            return
        self.overlap[:] = 0
        self.segments[:] = 0
        self.current = 0
        self.input_buffer[:] = 0
        self.pre_multiplied[:] = 0
        self.conv[:] = 0
        self.input_buffer_fill = 0


def compute_tail_block_size(head_len: int, response_len: int) -> int:
    """:514-526 in f32."""
    f = np.float32
    kn = (f(1.5) * f(head_len)) / (f(2.0) * np.log(f(2.0)))
    b = -kn + np.sqrt(kn * kn + f(response_len) * f(head_len), dtype=np.float32)
    b = max(f(b), f(head_len))
    return next_power_of_two(int(b))


class TwoStageNP:
    """:323-512"""

    @classmethod
    def init(cls, ir, block_size, max_response_length, forced_tail=0):
        ir = np.asarray(ir, dtype=np.float32)
        s = cls()
        s.head_block_size = block_size
        T = s.tail_block_size = forced_tail or compute_tail_block_size(block_size, max_response_length)
        if max_response_length < ir.size:
            raise Panic("max_response_length too small")
        L = max_response_length
        padded = np.zeros(L, np.float32)
        padded[: ir.size] = ir
        hl = min(L, T)
        s.head = FFTConvolverNP.init(padded[:hl], block_size, hl)
        if L > T:
            tl = min(L - T, T)
            s.tail0 = FFTConvolverNP.init(padded[T : T + tl], block_size, tl)
        else:
            s.tail0 = FFTConvolverNP()
        if L > 2 * T:
            tl = L - 2 * T
            s.tail = FFTConvolverNP.init(padded[2 * T : 2 * T + tl], T, tl)
        else:
            s.tail = FFTConvolverNP()
        s.tail_output0 = np.zeros(T, np.float32)
        s.tail_precalculated0 = np.zeros(T, np.float32)
        s.tail_output = np.zeros(T, np.float32)
        s.tail_precalculated = np.zeros(T, np.float32)
        s.tail_input = np.zeros(T, np.float32)
        s.tail_input_fill = 0
        s.precalculated_pos = 0
        return s

    def process(self, inp, out):
        if not len(inp) <= self.head_block_size:
            raise Panic("assert")
        H, T = self.head_block_size, self.tail_block_size
        self.head.process(inp, out)
        processed, ln = 0, len(inp)
        while processed < ln:
            n = min(ln - processed, H - (self.tail_input_fill % H))
            pp = self.precalculated_pos
            out[processed : processed + n] += self.tail_precalculated0[pp : pp + n]
            out[processed : processed + n] += self.tail_precalculated[pp : pp + n]
            self.precalculated_pos += n
            self.tail_input[self.tail_input_fill : self.tail_input_fill + n] = inp[processed : processed + n]
            self.tail_input_fill += n
            if self.tail_input_fill % H == 0:
                off = self.tail_input_fill - H
                self.tail0.process(self.tail_input[off : off + H], self.tail_output0[off : off + H])
                if self.tail_input_fill == T:
                    self.tail_precalculated0, self.tail_output0 = self.tail_output0, self.tail_precalculated0
            if self.tail_input_fill == T:
                self.tail_precalculated, self.tail_output = self.tail_output, self.tail_precalculated
                self.tail.process(self.tail_input, self.tail_output)
            if self.tail_input_fill == T:
                self.tail_input_fill = 0
                self.precalculated_pos = 0
            processed += n


def truth_f64(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """f64 linear convolution truncated to len(x) (FFT-based, for long signals)."""
    n = len(x) + len(h) - 1
    nfft = 1 << (n - 1).bit_length()
    y = np.fft.irfft(np.fft.rfft(x.astype(np.float64), nfft) * np.fft.rfft(h.astype(np.float64), nfft), nfft)
    return y[: len(x)]
