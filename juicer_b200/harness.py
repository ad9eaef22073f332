"""The callers either side of the decode path, host side only:

* HTK parameter-file reader — what Tracter's ``HTKSource`` feeds ``DecoderSingleTest`` with
  (12-byte big-endian header ``nSamples:i32 sampPeriod:i32 sampSize:i16 parmKind:i16`` + big-endian
  float32 frames; out of tree in the reference, format per the HTK book);
* extended file names ``name=file[s,e]`` exactly as ``DecoderSingleTest::configure`` carves them up
  (``src/DecoderSingleTest.cpp:118-157``) and the frame range ``decodeUtterance`` then feeds the
  decoder (``:262-296``, including its 20-frame pre-read, which ignores the end frame);
* word results from the decoder's word-boundary chain, as ``extractResultsFromHypWordMode``
  builds them (``src/DecoderSingleTest.cpp:402-469``: per-word scores are float32 differences of
  the cumulative ones, start time = previous end time, optional removal of sentence marks);
* the five output formats of ``DecoderBatchTest::outputResult`` (``src/DecoderBatchTest.cpp:264-459``):
  ref, trans, mlf, xmlf, verbose — byte for byte, including the time-stamp arithmetic of xmlf.

Nothing here touches the GPU; it turns files into the arrays ``WFSTDecoderLite.decode_batch`` takes and
its results into the text the reference prints.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

PREREAD = 20          # src/DecoderSingleTest.cpp:268


class HarnessError(ValueError):
    """The reference calls Torch3 error() (message + exit) in these cases."""


# ---------------------------------------------------------------------------------------
# input side
# ---------------------------------------------------------------------------------------
@dataclass
class ExtendedName:
    test_name: str            # symbolic name used in the outputs (DecoderSingleTest::getTestFName)
    data_file: str            # file actually read
    start: int = -1           # extStartFrame, -1 = none
    end: int = -1             # extEndFrame (inclusive), -1 = none


def parse_extended_filename(s: str) -> ExtendedName:
    """``name=file[s,e]`` / ``name=file`` / ``file`` — src/DecoderSingleTest.cpp:118-165, same checks."""
    if "=" not in s:
        return ExtendedName(s, s)
    name, _, rest = s.partition("=")
    if rest == "":
        raise HarnessError("DST::configure - error isolating real filename from extended filename")
    if "[" not in rest:
        return ExtendedName(name, rest)
    data, _, seg = rest.partition("[")
    if data == "":
        raise HarnessError("DST::configure - error isolating real filename from extended filename")
    first, sep, tail = seg.partition(",")
    try:
        start = int(first.strip().split()[0]) if first.strip() else None
    except ValueError:
        start = None
    if start is None:
        raise HarnessError("DST::configure - error extracting start frame from extended filename")
    if not sep or "]" not in tail:
        raise HarnessError("DST::configure - error isolating end frame from extended filename")
    try:
        end = int(tail.partition("]")[0].strip().split()[0])
    except (ValueError, IndexError):
        raise HarnessError("DST::configure - error extracting end frame from extended filename")
    if start < 0:
        raise HarnessError("DST::configure - extStartFrame < 0")
    if end <= 0:
        raise HarnessError("DST::configure - extEndFrame <= 0")
    if start >= end:
        raise HarnessError("DST::configure - extStartFrame >= extEndFrame")
    return ExtendedName(name, data, start, end)


def read_htk(path: str) -> Tuple[np.ndarray, int, int]:
    """Reads an (uncompressed) HTK parameter file: returns (frames float32 [n, dim], sampPeriod in
    100 ns units, parmKind).  Compressed (_C) and CRC-checked (_K) files are refused loudly."""
    with open(path, "rb") as f:
        hdr = f.read(12)
        if len(hdr) != 12:
            raise HarnessError(f"{path}: truncated HTK header")
        n, period, size, kind = struct.unpack(">iihh", hdr)
        if n < 0 or size <= 0 or size % 4:
            raise HarnessError(f"{path}: bad HTK header (nSamples={n}, sampSize={size})")
        if kind & 0o2000:
            raise HarnessError(f"{path}: compressed HTK files (_C) are not supported")
        body = f.read(n * size + (2 if kind & 0o10000 else 0))
    if len(body) < n * size:
        raise HarnessError(f"{path}: expected {n * size} data bytes, found {len(body)}")
    x = np.frombuffer(body[: n * size], dtype=">f4").astype(np.float32).reshape(n, size // 4)
    return x, period, kind & 0xffff


def write_htk(path: str, x: np.ndarray, samp_period: int = 100000, parm_kind: int = 6 | 0o400 | 0o1000) -> None:
    """Writes float32 frames as an HTK parameter file (default kind MFCC_D_A, 10 ms)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack(">iihh", x.shape[0], samp_period, x.shape[1] * 4, parm_kind))
        f.write(x.astype(">f4").tobytes())


def frames_decoded(n_file_frames: int, start: int = -1, end: int = -1) -> Tuple[int, int]:
    """(first frame, number of frames) DecoderSingleTest::decodeUtterance feeds the decoder for a
    file of n_file_frames frames and an optional [start, end] segment (src/DecoderSingleTest.cpp:
    262-296).  The end frame is inclusive — and the 20-frame pre-read (:272-275) does not look at it,
    so a segment shorter than 20 frames still decodes 20 (if the file has them)."""
    offset = 0 if start < 0 else start
    n_data = max(0, min(PREREAD, n_file_frames - offset))
    n = 0
    while n_data > 0:
        n += 1                                                     # processFrame(.., nFrames++, nData)
        nxt = n + offset + n_data - 1
        if (end < 0 or nxt - 1 < end) and nxt < n_file_frames:
            pass                                                   # one more frame fetched, nData unchanged
        else:
            n_data -= 1
    return offset, n


def load_utterance(spec: str, expected_dim: Optional[int] = None) -> Tuple[ExtendedName, np.ndarray, int]:
    """Extended name -> (parsed name, frames the reference would decode, sampPeriod)."""
    ext = parse_extended_filename(spec)
    x, period, _kind = read_htk(ext.data_file)
    if expected_dim is not None and x.shape[1] != expected_dim:
        raise HarnessError(f"{ext.data_file}: vector size {x.shape[1]} != expected {expected_dim}")
    first, n = frames_decoded(x.shape[0], ext.start, ext.end)
    return ext, np.ascontiguousarray(x[first:first + n]), period


# ---------------------------------------------------------------------------------------
# result side
# ---------------------------------------------------------------------------------------
@dataclass
class ResultWord:                     # DSTResultWord
    index: int                        # vocabulary index = output label - 1
    start_time: int
    end_time: int
    acoustic_score: np.float32
    lm_score: np.float32


def extract_result_words(labels: Sequence[int], times: Sequence[int], ac: Sequence[float], lm: Sequence[float],
                         sent_start_index: int = -1, sent_end_index: int = -1,
                         remove_sent_marks: bool = False) -> List[ResultWord]:
    """extractResultsFromHypWordMode (src/DecoderSingleTest.cpp:402-469).  labels/times/ac/lm are the
    word-boundary records oldest first (JgpuResult.words: label = hist->state, cumulative scores)."""
    keep = [i for i, l in enumerate(labels)
            if not remove_sent_marks or ((l - 1) != sent_start_index and (l - 1) != sent_end_index)]
    out = [ResultWord(int(labels[i]) - 1, 0, int(times[i]), np.float32(ac[i]), np.float32(lm[i])) for i in keep]
    # the reference walks newest -> oldest and, when it fills word w, differences word w+1 against it
    for w in range(len(out) - 2, -1, -1):
        out[w + 1].start_time = out[w].end_time
        out[w + 1].acoustic_score = np.float32(out[w + 1].acoustic_score - out[w].acoustic_score)
        out[w + 1].lm_score = np.float32(out[w + 1].lm_score - out[w].lm_score)
    if out:
        out[0].start_time = 0
    return out


def _rec_name(test_name: str) -> str:
    base = test_name.rsplit("/", 1)[-1]
    dot = base.rfind(".")
    return base[:dot] if dot >= 0 else base


def format_result(fmt: str, test_name: str, words: Sequence[str], result: Sequence[ResultWord], n_frames: int,
                  frames_per_sec: int = 100, frame_time0_ns: int = 0,
                  expected: Optional[Sequence[int]] = None) -> str:
    """One utterance in one of DecoderBatchTest's output formats (src/DecoderBatchTest.cpp:343-436);
    `words` maps vocabulary index -> word string.  An utterance without a result prints what the
    reference prints for nResultWords == 0 (an empty ref line, an empty MLF entry, ...)."""
    toks = [words[r.index] for r in result]
    if fmt == "ref":
        return "".join(t + " " for t in toks) + "\n"
    if fmt == "trans":
        return "".join(t + " " for t in toks) + f"(trans-{len(toks)})\n"
    if fmt in ("mlf", "xmlf"):
        s = f"\"*/{_rec_name(test_name)}.rec\"\n"
        if fmt == "mlf":
            s += "".join(t + "\n" for t in toks)
        else:
            unit = np.float32(1.0e7) / np.float32(frames_per_sec)          # (real)1.0e7 / (real)framesPerSec
            offset = float(frame_time0_ns) / 100
            for r, t in zip(result, toks):
                st = float(unit * np.float32(r.start_time))
                if st > 0:
                    st += float(unit)
                et = float(unit * np.float32(r.end_time))
                if et > 0:
                    et += float(unit)
                score = float(np.float32(r.acoustic_score + r.lm_score))
                s += "%.0f %.0f %s %f\n" % (st + offset, et + offset, t, score)
        return s + ".\n"
    if fmt == "verbose":
        s = test_name + "\n"
        if expected is not None:
            s += "\tExpected :  " + "".join(("<OOV> " if e < 0 else words[e] + " ") for e in expected) + "\n"
        s += "\tActual :    " + "".join(t + " " for t in toks)
        s += "  [ " + "".join(f"{r.end_time + 1} " for r in result) + f"({n_frames}) ]\n"
        return s
    raise HarnessError(f"unknown output format {fmt!r} (ref, trans, mlf, xmlf, verbose)")


def mlf_header() -> str:
    return "#!MLF!#\n"                 # DecoderBatchTest::openOutputFile writes it for mlf / xmlf


def decode_files(decoder, specs: Sequence[str], words: Sequence[str], fmt: str = "mlf", expected_dim: Optional[int] = None,
                 frames_per_sec: int = 100, remove_sent_marks: bool = False, sent_start_index: int = -1,
                 sent_end_index: int = -1) -> str:
    """DecoderBatchTest::run over a list of (extended) HTK file names with the GPU decoder: one
    lock-step batch instead of a per-file loop; returns the text of the output file."""
    loaded = [load_utterance(s, expected_dim) for s in specs]
    results = decoder.decode_batch([x for _e, x, _p in loaded])
    out = mlf_header() if fmt in ("mlf", "xmlf") else ""
    for (ext, x, _p), r in zip(loaded, results):
        rw: List[ResultWord] = []
        if r.status > 0:
            rw = extract_result_words(r.labels, r.times, [w["ac"] for w in r.words], [w["lm"] for w in r.words],
                                      sent_start_index, sent_end_index, remove_sent_marks)
        out += format_result(fmt, ext.test_name, words, rw, x.shape[0], frames_per_sec)
    return out
