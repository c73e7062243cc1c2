"""Host-side mirror of JointAnnotatorCaller (CORE/genotyping/JointAnnotatorCaller.scala) over the
C ABI (avo_joint_counts / avo_joint), plus the sample-sharded multi-process form.

    JointAnnotatorCaller.apply(genotypes)   -> JointAnnotatorCaller.scala:52-64
    annotateSite                            -> :74-106

The reference groups a GenotypeRDD into VariantContexts (one per variant, the genotypes of all
samples); here the grouping is the (samples x sites) matrix the library consumes, sample-major.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .batch import _device_empty, _ptr
from .records import ALT, NO_CALL, OTHER_ALT, REF, Genotype, Variant


def _is_dev(a):
    return a is not None and not isinstance(a, np.ndarray)


def _joint_in(n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl, present=None, n_nocall=None,
              totals=None):
    s = L.JointInStruct()
    s.n_sites, s.n_samples, s.n_states = n_sites, n_samples, n_states
    s.mem = L.MEM_DEVICE if _is_dev(n_alt) else L.MEM_HOST
    s.present, s.n_alt, s.n_ref, s.n_other = _ptr(present), _ptr(n_alt), _ptr(n_ref), _ptr(n_other)
    s.n_nocall, s.gl = _ptr(n_nocall), _ptr(gl)
    if totals is not None:
        s.alt_alleles, s.called_alleles = _ptr(totals[0]), _ptr(totals[1])
    return s


def joint_counts(ctx, n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl, present=None,
                 n_nocall=None, out=None):
    """avo_joint_counts: (alt alleles, called alleles) per site over these samples.  `out`: one int32
    buffer of 2 x n_sites entries that receives both (a sharded cohort all-reduces it as ONE message)."""
    dev = _is_dev(n_alt)
    if out is None:
        out = _device_empty(2 * n_sites, np.int32, "cuda:%d" % ctx.device) if dev else np.zeros(max(2 * n_sites, 2), np.int32)
    alt, called = out[:n_sites], out[n_sites:2 * n_sites]
    s = _joint_in(n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl, present, n_nocall)
    ctx._check(ctx.lib.avo_joint_counts(ctx.h, C.byref(s), _ptr(alt), _ptr(called)))
    return alt[:n_sites], called[:n_sites]


def joint_call(ctx, n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl, present=None,
               n_nocall=None, totals=None) -> Dict[str, object]:
    """avo_joint on (samples x sites) columns (numpy or torch CUDA tensors, sample-major).
    `totals` = (alt_alleles, called_alleles) of the whole cohort when these samples are a shard."""
    dev = _is_dev(n_alt)
    s = _joint_in(n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl, present, n_nocall, totals)
    o = L.JointOutStruct()
    o.n_sites, o.n_samples, o.n_states = n_sites, n_samples, n_states
    o.mem = s.mem
    out = {}
    for name, dt, kind in L.JOINT_OUT_COLUMNS:
        n = {"N": n_sites, "SN": n_sites * n_samples, "SNG": n_sites * n_samples * n_states}[kind]
        a = _device_empty(n, np.dtype(dt).type, "cuda:%d" % ctx.device) if dev else np.zeros(max(n, 1), dt)
        out[name] = a
        setattr(o, name, _ptr(a))
    ctx._check(ctx.lib.avo_joint(ctx.h, C.byref(s), C.byref(o)))
    shaped = {}
    for name, dt, kind in L.JOINT_OUT_COLUMNS:
        a = out[name]
        if kind == "N":
            shaped[name] = a[:n_sites]
        elif kind == "SN":
            shaped[name] = a[:n_sites * n_samples].reshape(n_samples, n_sites)
        else:
            shaped[name] = a[:n_sites * n_samples * n_states].reshape(n_samples, n_sites, n_states)
    return shaped


def joint_depths(ctx, n_sites, n_samples, read_depth=None, ref_read_depth=None, sb=None, present=None
                 ) -> Dict[str, object]:
    """avo_joint_depths on (samples x sites) columns (numpy or torch CUDA, sample-major; sb is [S, N, 4]):
    the per-site VariantAnnotation depths of calculateAnnotations and the `annotated` flag (-1 = unset)."""
    given = [a for a in (read_depth, ref_read_depth, sb, present) if a is not None]
    dev = bool(given) and _is_dev(given[0])
    i = L.DepthsInStruct()
    i.n_sites, i.n_samples, i.mem = n_sites, n_samples, L.MEM_DEVICE if dev else L.MEM_HOST
    keep = []
    for name, a, dt in (("present", present, np.uint8), ("read_depth", read_depth, np.int32),
                        ("ref_read_depth", ref_read_depth, np.int32), ("sb", sb, np.int32)):
        if a is None:
            continue
        if not dev:
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
        setattr(i, name, _ptr(a))
    o = L.SiteDepthsStruct()
    o.n_sites, o.mem = n_sites, i.mem
    out = {}
    for name, dt in [("annotated", np.uint8)] + [(c, np.int32) for c in L.SITE_DEPTH_COLUMNS]:
        a = _device_empty(n_sites, dt, "cuda:%d" % ctx.device) if dev else np.zeros(max(n_sites, 1), dt)
        out[name] = a
        setattr(o, name, _ptr(a))
    ctx._check(ctx.lib.avo_joint_depths(ctx.h, C.byref(i), C.byref(o)))
    return {k: v[:n_sites] for k, v in out.items()}


def reduce_site_depths(local: Dict[str, object], n_genotypes, group=None) -> Dict[str, object]:
    """The exchange step of the depth roll-up when the cohort is sharded by sample: `local`
    is this rank's joint_depths output (-1 = unset), `n_genotypes` its genotypes per site.  One all-reduce
    (sum) of 13 int32 vectors -- the six sums, how many ranks have each field, the genotype counts -- gives
    every rank the cohort's columns: a field is unset only if no rank has it (VariantSummary.merge is
    associative), a site is annotated when the cohort has more than one genotype there.
    CUDA tensors (NCCL) or numpy arrays (gloo)."""
    import torch
    import torch.distributed as dist
    dev = _is_dev(local[L.SITE_DEPTH_COLUMNS[0]])
    t = lambda a: a if dev else torch.from_numpy(np.ascontiguousarray(a))
    cols = [t(local[c]).to(torch.int32) for c in L.SITE_DEPTH_COLUMNS]
    pack = torch.stack([c.clamp(min=0) for c in cols] + [(c >= 0).to(torch.int32) for c in cols] +
                       [t(n_genotypes).to(torch.int32)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    out = {c: torch.where(pack[6 + k] > 0, pack[k], torch.full_like(pack[k], -1)) for k, c in enumerate(L.SITE_DEPTH_COLUMNS)}
    out["annotated"] = (pack[12] > 1).to(torch.uint8)
    return out if dev else {k: v.numpy() for k, v in out.items()}


def joint_depths_sharded(ctx, n_sites, n_samples, read_depth=None, ref_read_depth=None, sb=None, present=None,
                         group=None) -> Dict[str, object]:
    """avo_joint_depths on this rank's samples, then reduce_site_depths over the ranks of `group`."""
    local = joint_depths(ctx, n_sites, n_samples, read_depth, ref_read_depth, sb, present)
    if present is None:
        import torch
        cnt = torch.full((n_sites,), n_samples, dtype=torch.int32, device=local["annotated"].device) \
            if _is_dev(local["annotated"]) else np.full(n_sites, n_samples, np.int32)
    else:
        cnt = present.reshape(n_samples, n_sites).sum(0)
    return reduce_site_depths(local, cnt, group)


MINUS_TEN_DIV_LOG10 = -10.0 / np.log(10.0)


def joint_call_sharded(ctx, cols, group=None) -> Dict[str, object]:
    """Cohort sharded by sample over the ranks of `group`: local allele counts,
    all-reduce(sum) of the two int32 vectors, local re-call against the cohort totals, all-reduce
    (sum) of the per-site posterior(0) sums, quality = -10/ln10 * sum.  `cols` holds this rank's
    samples: n_sites, n_samples, n_states, n_alt, n_ref, n_other, gl [, present, n_nocall] as CUDA
    tensors (NCCL) or numpy arrays (gloo); every rank must hold the same sites in the same order."""
    import torch
    import torch.distributed as dist
    n = cols["n_sites"]
    dev = _is_dev(cols["n_alt"])
    both = _device_empty(2 * n, np.int32, "cuda:%d" % ctx.device) if dev else np.zeros(max(2 * n, 2), np.int32)
    joint_counts(ctx, out=both, **cols)
    t_both = both if dev else torch.from_numpy(both)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t_both, op=dist.ReduceOp.SUM, group=group)     # both count vectors, one message
    totals = (t_both[:n], t_both[n:2 * n]) if dev else (both[:n], both[n:2 * n])
    out = joint_call(ctx, totals=totals, **cols)
    p0 = out["post0_sum"] if dev else torch.from_numpy(out["post0_sum"])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(p0, op=dist.ReduceOp.SUM, group=group)
    out["quality"] = p0 * MINUS_TEN_DIV_LOG10 if dev else p0.numpy() * MINUS_TEN_DIV_LOG10
    out["alt_alleles"], out["called_alleles"] = totals
    return out


@dataclass
class VariantSummary:
    """CORE/genotyping/VariantSummary.scala:39-146: genotype depths rolled up to the site (None = unset)."""
    readDepth: Optional[int] = None
    referenceReadDepth: Optional[int] = None
    forwardReadDepth: Optional[int] = None
    reverseReadDepth: Optional[int] = None
    forwardReferenceReadDepth: Optional[int] = None
    reverseReferenceReadDepth: Optional[int] = None

    @staticmethod
    def of(g: Genotype) -> "VariantSummary":                                       # :39-70
        sb = list(g.strandBiasComponents or [])
        if not sb:
            return VariantSummary(g.readDepth, g.referenceReadDepth)
        if len(sb) != 4:
            raise ValueError("Genotype (%r) has an illegal number of strand bias components." % (g,))
        rfs, rrs, afs, ars = (int(x) for x in sb)          # 0,1 ref; 2,3 alt; 0,2 forward; 1,3 reverse
        return VariantSummary(g.readDepth, g.referenceReadDepth, rfs + afs, rrs + ars, rfs, rrs)

    def merge(self, o: "VariantSummary") -> "VariantSummary":                      # :95-119
        def add(a, b):
            return a + b if a is not None and b is not None else a if a is not None else b
        return VariantSummary(*(add(getattr(self, f.name), getattr(o, f.name)) for f in dataclasses.fields(self)))

    def to_annotation(self, existing: Optional[Dict[str, int]] = None) -> Dict[str, int]:      # :126-145
        """VariantAnnotation fields: the variant's current annotation with every set field replaced."""
        names = {"readDepth": "readDepth", "referenceReadDepth": "referenceReadDepth",
                 "forwardReadDepth": "forwardReadDepth", "reverseReadDepth": "reverseReadDepth",
                 "forwardReferenceReadDepth": "referenceForwardReadDepth",
                 "reverseReferenceReadDepth": "referenceReverseReadDepth"}
        out = dict(existing or {})
        for mine, theirs in names.items():
            if getattr(self, mine) is not None:
                out[theirs] = int(getattr(self, mine))
        return out


def calculate_annotations(genotypes: Sequence[Genotype], existing: Optional[Dict[str, int]] = None) -> Dict[str, int]:
    """JointAnnotatorCaller.calculateAnnotations (:141-147): site.genotypes.map(VariantSummary(_)).reduce(_.merge(_))."""
    acc = None
    for g in genotypes:
        v = VariantSummary.of(g)
        acc = v if acc is None else acc.merge(v)
    return (acc or VariantSummary()).to_annotation(existing)


@dataclass
class VariantContext:
    """One squared-off site: the variant (with its joint quality) and the genotypes of all samples."""
    variant: Variant
    quality: float
    genotypes: List[Genotype]
    minorAlleleFrequency: float
    annotation: Optional[Dict[str, int]] = None       # the rolled-up VariantAnnotation depths (VariantSummary)


class JointAnnotatorCaller:
    """CORE/genotyping/JointAnnotatorCaller.scala: recall and site quality (avo_joint) and the rolled-up
    depth annotation of :141-147 (avo_joint_depths) on the GPU."""

    @staticmethod
    def apply(genotypes: Sequence[Genotype], ctx=None) -> List[VariantContext]:
        from .genotyping import default_context
        ctx = ctx or default_context()
        samples = sorted({g.sampleId for g in genotypes}, key=lambda x: (x is None, x))
        s_of = {s: i for i, s in enumerate(samples)}
        keys: Dict[tuple, int] = {}
        for g in genotypes:                       # toVariantContexts: group by variant
            keys.setdefault(g.variant.as_tuple(), len(keys))
        n, S = len(keys), len(samples)
        if n == 0:
            return []
        ns = max(max(len(g.genotypeLikelihoods) for g in genotypes), 2)
        if ns > L.MAX_PLOIDY + 1:
            raise L.AvocadoError(L.AVO_E_UNSUPPORTED, "more than %d genotype states" % (L.MAX_PLOIDY + 1))
        present = np.zeros((S, n), np.uint8)
        cnt = {a: np.zeros((S, n), np.int32) for a in (ALT, REF, OTHER_ALT, NO_CALL)}
        gl = np.zeros((S, n, ns), np.float64)
        where: Dict[tuple, Genotype] = {}
        for g in genotypes:
            s, i = s_of[g.sampleId], keys[g.variant.as_tuple()]
            if present[s, i]:
                raise ValueError("two genotypes of sample %r at %r" % (g.sampleId, g.variant))
            present[s, i] = 1
            for a in g.alleles:
                cnt[a][s, i] += 1
            gl[s, i, :len(g.genotypeLikelihoods)] = g.genotypeLikelihoods
            where[(s, i)] = g
        out = joint_call(ctx, n, S, ns, cnt[ALT].ravel(), cnt[REF].ravel(), cnt[OTHER_ALT].ravel(),
                         gl.ravel(), present.ravel(), cnt[NO_CALL].ravel())
        # calculateAnnotations (:141-147) for every site: avo_joint_depths
        dp, refdp, sb = -np.ones((S, n), np.int32), -np.ones((S, n), np.int32), -np.ones((S, n, 4), np.int32)
        for (s, i), g in where.items():
            dp[s, i] = -1 if g.readDepth is None else g.readDepth
            refdp[s, i] = -1 if g.referenceReadDepth is None else g.referenceReadDepth
            comp = list(g.strandBiasComponents or [])
            if comp:
                if len(comp) != 4:                                          # VariantSummary.scala:49-50 (require)
                    raise ValueError("Genotype (%r) has an illegal number of strand bias components." % (g,))
                sb[s, i] = comp
        depths = joint_depths(ctx, n, S, dp.ravel(), refdp.ravel(), sb.ravel(), present.ravel())
        ann_names = dict(zip(L.SITE_DEPTH_COLUMNS, ("readDepth", "referenceReadDepth", "forwardReadDepth",
                                                    "reverseReadDepth", "referenceForwardReadDepth",
                                                    "referenceReverseReadDepth")))
        sites: List[VariantContext] = []
        for key, i in keys.items():
            if not out["site_emitted"][i]:
                continue                                                    # :81-83
            gts = []
            for s in range(S):
                g = where.get((s, i))
                if g is None:
                    continue
                fl = int(out["flags"][s, i])
                k = len(g.genotypeLikelihoods)
                new = dataclasses.replace(g, nonReferenceLikelihoods=[])         # :231, :247, :259
                if fl & L.JOINT_POSTERIORS:
                    new.genotypePosteriors = [float(x) for x in out["posteriors"][s, i, :k]]
                if fl & L.JOINT_PRIORS:
                    new.genotypePriors = [float(x) for x in out["priors"][s, i, :k]]
                if fl & L.JOINT_RECALLED:
                    new.alleles = [ALT] * int(out["n_alt"][s, i]) + [REF] * int(out["n_ref"][s, i])
                    new.genotypeQuality = int(out["gq"][s, i])
                gts.append(new)
            # :85-90: "if we have more than one sample, roll up annotations"
            ann = None
            if depths["annotated"][i]:
                ann = {ann_names[c]: int(depths[c][i]) for c in L.SITE_DEPTH_COLUMNS if depths[c][i] >= 0}
            sites.append(VariantContext(Variant(*key), float(out["quality"][i]), gts, float(out["maf"][i]), ann))
        return sites
