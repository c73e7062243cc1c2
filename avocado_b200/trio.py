"""Host-side mirror of TrioCaller (CORE/genotyping/TrioCaller.scala:60-221) over the C ABI
(avo_trio_call): the per-site parent / child consistency and phasing pass of the trio workflow
(CLI/TrioGenotyper.scala)."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .batch import _device_empty, _ptr
from .joint import VariantContext
from .records import ALT, NO_CALL, OTHER_ALT, REF, Genotype, Variant


def trio_call_packed(ctx, n_sites, n_alt, n_ref, n_other, present=None, n_nocall=None) -> Dict[str, object]:
    """avo_trio_call on the (3 x sites) matrix rows (first parent, second parent, child), numpy or
    torch CUDA, sample-major.  Returns kept / child_action / filled per site."""
    dev = not isinstance(n_alt, np.ndarray)
    t = L.TrioStruct()
    t.n_sites = n_sites
    t.mem = L.MEM_DEVICE if dev else L.MEM_HOST
    t.present, t.n_alt, t.n_ref, t.n_other, t.n_nocall = (_ptr(a) for a in (present, n_alt, n_ref, n_other, n_nocall))
    out = {}
    for name in ("kept", "child_action", "filled"):
        a = _device_empty(n_sites, np.uint8, "cuda:%d" % ctx.device) if dev else np.zeros(max(n_sites, 1), np.uint8)
        out[name] = a
        setattr(t, name, _ptr(a))
    ctx._check(ctx.lib.avo_trio_call(ctx.h, C.byref(t)))
    return {k: v[:n_sites] for k, v in out.items()}


class TrioCaller:
    @staticmethod
    def apply(genotypes: Sequence[Genotype], firstParentId: str, secondParentId: str, childId: str, ctx=None
              ) -> List[VariantContext]:
        """TrioCaller.apply (:75-97) on genotypes grouped by variant."""
        from .genotyping import default_context
        ctx = ctx or default_context()
        ids = [firstParentId, secondParentId, childId]
        keys: Dict[tuple, int] = {}
        for g in genotypes:
            keys.setdefault(g.variant.as_tuple(), len(keys))
        n = len(keys)
        if n == 0:
            return []
        present = np.zeros((3, n), np.uint8)
        cnt = {a: np.zeros((3, n), np.int32) for a in (ALT, REF, OTHER_ALT, NO_CALL)}
        where, others = {}, {}
        for g in genotypes:
            i = keys[g.variant.as_tuple()]
            if g.sampleId in ids and not present[ids.index(g.sampleId), i]:      # genotypes.find: the first one
                s = ids.index(g.sampleId)
                present[s, i] = 1
                for a in g.alleles:
                    cnt[a][s, i] += 1
                where[(s, i)] = g
        # filterRef looks at every genotype of the site (:104-111); samples outside the trio are not emitted
        out = trio_call_packed(ctx, n, cnt[ALT].ravel(), cnt[REF].ravel(), cnt[OTHER_ALT].ravel(), present.ravel(),
                               cnt[NO_CALL].ravel())
        res = []
        for key, i in keys.items():
            if not out["kept"][i]:
                continue
            v = Variant(*key)
            gts = []
            for s in range(3):
                g = where.get((s, i))
                if g is None:                                                       # makeNoCall (:119-127)
                    g = Genotype(variant=v, contigName=v.contigName, start=v.start, end=v.end, sampleId=ids[s],
                                 alleles=[NO_CALL, NO_CALL], genotypeQuality=None, genotypeLikelihoods=[],
                                 nonReferenceLikelihoods=[], strandBiasComponents=[], readDepth=0,
                                 referenceReadDepth=0, alternateReadDepth=0, rmsMapQ=float("nan"),
                                 fisherStrandBiasPValue=float("nan"))
                elif s == 2:
                    act = int(out["child_action"][i])
                    if act == L.TRIO_NO_CALL:
                        g = dataclasses.replace(g, alleles=[NO_CALL, NO_CALL], genotypeQuality=None)
                    elif act == L.TRIO_PHASED:
                        g = dataclasses.replace(g, phased=True)
                    elif act == L.TRIO_PHASED_ALT_REF:
                        g = dataclasses.replace(g, phased=True, alleles=[ALT, REF])
                    elif act == L.TRIO_PHASED_REF_ALT:
                        g = dataclasses.replace(g, phased=True, alleles=[REF, ALT])
                gts.append(g)
            res.append(VariantContext(v, float("nan"), gts, float("nan")))
        return res
