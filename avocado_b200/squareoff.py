"""Host-side mirror of SquareOffReferenceModel (CORE/genotyping/SquareOffReferenceModel.scala) over
the C ABI (avo_square_off): the step between per-sample gVCF calling and joint calling
(docs/workflows/joint.rst).

    extractVariants (:125-166)   sites where some sample carries an ALT allele, alleles right-trimmed,
                                 de-duplicated -- host side (a few 1e4 rows per sample);
    squareOffSite (:176-244)     every site x every sample: the sample's own genotype for exactly
                                 that variant, else the first overlapping genotype with its
                                 non-reference likelihoods -- the merge-join, on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib as L
from .batch import SiteTable, _device_empty, _ptr, sites_from_variants
from .records import Variant


def trim_right(ref: str, alt: str) -> int:
    """SquareOffReferenceModel.trimRight (:95-117): common suffix length, never consuming a whole allele."""
    if not ref or not alt:
        return 0
    i, j, trimmed = len(ref), len(alt), 0
    while True:
        i -= 1
        j -= 1
        if ref[i] != alt[j] or not (i > 0 and j > 0):
            return trimmed
        trimmed += 1


def extract_variants(variants: Sequence[Variant], has_alt_call: Sequence[bool]) -> List[Variant]:
    """extractVariants (:125-166): variants of the genotypes that carry an ALT allele, right-trimmed
    to a canonical form, de-duplicated (order of first appearance)."""
    seen, out = set(), []
    for v, called in zip(variants, has_alt_call):
        if not called or v.alternateAllele is None:
            continue
        t = trim_right(v.referenceAllele, v.alternateAllele)
        nv = Variant(v.contigName, v.start, v.end - t, v.referenceAllele[:len(v.referenceAllele) - t],
                     v.alternateAllele[:len(v.alternateAllele) - t]) if t else v
        if nv.as_tuple() not in seen:
            seen.add(nv.as_tuple())
            out.append(nv)
    return out


def _results_struct(cols, n, ns, mem):
    r = L.SiteResultsStruct()
    r.n_sites, r.n_states, r.mem = n, ns, mem
    keep = []
    for name, dt, w in L.RESULT_COLUMNS:
        a = cols.get(name)
        if a is None:
            continue
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, np.dtype(dt))
            keep.append(a)
        setattr(r, name, _ptr(a))
    return r, keep


def square_off_packed(ctx, cohort: SiteTable, sample_sites: SiteTable, sample_cols: Dict[str, object],
                      ref_start, ref_cols: Dict[str, object], n_states: int, ref_len=None,
                      ref_max_len: int = 0) -> Dict[str, object]:
    """avo_square_off for one sample: `sample_sites` / `sample_cols` are the table and variant rows
    given to / returned by call_all_sites_packed, `ref_start` / `ref_cols` its reference-model rows.
    `ref_len` (optional): the rows are blocks covering [start, start + len) as in a banded gVCF;
    `ref_max_len` bounds the lengths of device-resident blocks.
    Returns the SQUARE_OUT_COLUMNS (one row per cohort site)."""
    dev = cohort.on_device
    mem = L.MEM_DEVICE if dev else L.MEM_HOST
    cs, vs = cohort.as_struct(), sample_sites.as_struct()
    vr, k1 = _results_struct(sample_cols, sample_sites.n, n_states, mem)
    n_r = len(ref_start)
    rr, k2 = _results_struct(ref_cols, n_r, n_states, mem)
    if not dev:
        ref_start = np.ascontiguousarray(ref_start, np.int32)
        if ref_len is not None:
            ref_len = np.ascontiguousarray(ref_len, np.int32)
    sr = L.SampleRowsStruct()
    sr.variants, sr.variant_rows = C.pointer(vs), C.pointer(vr)
    sr.n_ref_model, sr.ref_model_start, sr.ref_model_rows = n_r, _ptr(ref_start), C.pointer(rr)
    if ref_len is not None:
        sr.ref_model_len, sr.ref_model_max_len = _ptr(ref_len), int(ref_max_len)
    out = L.SquareOutStruct()
    out.n_sites, out.n_states, out.mem = cohort.n, n_states, mem
    res = {}
    for name, dt, w in L.SQUARE_OUT_COLUMNS:
        n = cohort.n * (n_states if w == "ns" else 1)
        a = _device_empty(n, np.dtype(dt).type, "cuda:%d" % ctx.device) if dev else np.zeros(max(n, 1), dt)
        res[name] = a
        setattr(out, name, _ptr(a))
    ctx._check(ctx.lib.avo_square_off(ctx.h, C.byref(cs), C.byref(sr), C.byref(out)))
    shaped = {}
    for name, dt, w in L.SQUARE_OUT_COLUMNS:
        a = res[name][:cohort.n * (n_states if w == "ns" else 1)]
        shaped[name] = a.reshape(cohort.n, n_states) if w == "ns" else a
    return shaped
