"""ctypes binding of libavocado_b200.so (the C ABI in include/avocado_b200.h).

There is no fallback: if the shared library is missing the import fails loudly with the command
that builds it, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVO_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("AVO_LIB") or os.path.join(_HERE, "libavocado_b200.so")

AVO_OK = 0
AVO_E_INVALID = -1
AVO_E_CUDA = -2
AVO_E_CAPACITY = -3
AVO_E_EMPTY_SITES = -4
AVO_E_UNSORTED = -5
AVO_E_UNSUPPORTED = -6
AVO_E_NOMEM = -7
AVO_E_RETRY = -8
MEM_HOST, MEM_DEVICE = 0, 1
MAX_PLOIDY = 7

OP_MATCH, OP_MISMATCH, OP_INS, OP_DEL, OP_SOFTCLIP, OP_HARDCLIP = range(6)
READ_NEGATIVE_STRAND, READ_SKIP_CALL, READ_OPS_DROPPED = 1, 2, 4

STATUS_NAMES = {0: "AVO_OK", -1: "AVO_E_INVALID", -2: "AVO_E_CUDA", -3: "AVO_E_CAPACITY",
                -4: "AVO_E_EMPTY_SITES", -5: "AVO_E_UNSORTED", -6: "AVO_E_UNSUPPORTED",
                -7: "AVO_E_NOMEM", -8: "AVO_E_RETRY"}

P = C.c_void_p


class ReadBatchStruct(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_blocks", C.c_int64), ("n_ops", C.c_int64),
                ("n_refb", C.c_int64), ("contig_id", C.c_int32), ("mem", C.c_int32),
                ("stats_min_start", C.c_int32), ("stats_max_end", C.c_int32),
                ("stats_max_span", C.c_int32), ("stats_valid", C.c_int32),
                ("meta", P), ("payload", P), ("ops", P), ("refb", P), ("wire_meta", P), ("wire_ops", P),
                ("wire6_meta", P), ("wire6_oplen", P), ("wire6_opcode", P), ("wire6_exc", P), ("n_wire6_exc", C.c_int64), ("wire6_base", C.c_int32),
                ("reserved", C.c_int32), ("events", P)]


class ReadEventsStruct(C.Structure):
    """avo_read_events: discovery's candidate events of a batch (avo_derive_events)."""
    _fields_ = [("n_snv", C.c_int64), ("n_indel", C.c_int64), ("snv_pos", P), ("snv_pq", P), ("snv_first", P),
                ("indel_read", P), ("indel_op", P),
                ("stats_min_start", C.c_int32), ("stats_max_end", C.c_int32), ("stats_max_span", C.c_int32),
                ("stats_valid", C.c_int32)]


class SitesStruct(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_allele_nibbles", C.c_int64), ("mem", C.c_int32),
                ("reserved", C.c_int32), ("contig", P), ("start", P), ("ref_len", P), ("alt_len", P),
                ("allele_off", P), ("alleles4", P)]


class CnvStruct(C.Structure):
    _fields_ = [("n", C.c_int64), ("contig", P), ("start", P), ("end", P), ("copy_number", P)]


class ParamsStruct(C.Structure):
    _fields_ = [("ploidy", C.c_int32), ("score_all_sites", C.c_int32), ("max_quality", C.c_int32),
                ("max_mapq", C.c_int32), ("min_phred_discover", C.c_int32),
                ("min_obs_discover", C.c_int32), ("host_payload", C.c_int32), ("gather_threads", C.c_int32)]


class SitesOutStruct(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("allele_capacity", C.c_int64), ("n", C.c_int64),
                ("n_allele_nibbles", C.c_int64), ("mem", C.c_int32), ("reserved", C.c_int32),
                ("start", P), ("ref_len", P), ("alt_len", P), ("allele_off", P), ("alleles4", P),
                ("count", P)]


class ShardPlanStruct(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32),
                ("site_cap", C.c_int64), ("allele_cap", C.c_int64), ("row_cap", C.c_int64),
                ("n_states", C.c_int32), ("reserved", C.c_int32)]


RESULT_COLUMNS = [
    # name, dtype, per-site width ("ns" = n_states)
    ("emitted", "uint8", 1), ("allele_fwd", "int32", 1), ("other_fwd", "int32", 1),
    ("allele_cov", "int32", 1), ("other_cov", "int32", 1), ("total_cov", "int32", 1),
    ("square_mapq", "float64", 1), ("ref_ll", "float64", "ns"), ("allele_ll", "float64", "ns"),
    ("other_ll", "float64", "ns"), ("nonref_ll", "float64", "ns"), ("first_is_ref", "uint8", 1),
    ("copy_number", "int32", 1), ("n_alt", "int32", 1), ("n_ref", "int32", 1), ("n_other", "int32", 1),
    ("qual", "float64", 1), ("gq", "int32", 1), ("sb", "int32", 4), ("fisher", "float32", 1),
    ("rms_mapq", "float32", 1), ("gl", "float64", "ns"), ("nonref_gl", "float64", "ns"),
]


class SiteResultsStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_states", C.c_int32), ("mem", C.c_int32)] + \
               [(name, P) for name, _, _ in RESULT_COLUMNS]


class RefModelOutStruct(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("n", C.c_int64), ("start", P), ("rows", SiteResultsStruct)]


class TimingStruct(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_discover", C.c_float),
                ("ms_observe", C.c_float), ("ms_d2h", C.c_float), ("ms_discover_tiles", C.c_float),
                ("ms_call_tiles", C.c_float), ("n_launches", C.c_int32),
                ("n_tile_launches", C.c_int32), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("payload_in_place", C.c_int32), ("ms_call_pairs", C.c_float),
                ("gathered_blocks", C.c_int64)]


class TextReadsStruct(C.Structure):
    _fields_ = [("n", C.c_int64), ("start", P), ("end", P), ("mapq", P), ("rec_flags", P),
                ("cigar", P), ("cigar_off", P), ("md", P), ("md_off", P), ("seq", P), ("seq_off", P),
                ("qual", P), ("qual_off", P), ("keep", P), ("prefilter", P), ("narrow", C.c_int32), ("reserved", C.c_int32)]


class PrefilterStruct(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("keep_duplicates", C.c_int32), ("keep_non_primary", C.c_int32),
                ("contig_kept", C.c_int32)]


class TextOutStruct(C.Structure):
    _fields_ = [("cigar_bytes", C.c_int64), ("md_bytes", C.c_int64), ("seq_bytes", C.c_int64),
                ("start", P), ("end", P), ("mapq", P), ("rec_flags", P), ("cigar", P), ("cigar_off", P),
                ("md", P), ("md_off", P), ("seq", P), ("qual", P), ("seq_off", P)]


class PackedOutStruct(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_blocks", C.c_int64), ("n_ops", C.c_int64),
                ("n_refb", C.c_int64), ("meta", P), ("payload", P), ("ops", P),
                ("refb", P), ("source_index", P)]


FILTER_PARAM_FIELDS = [
    ("min_quality", C.c_int32), ("filter_ref_genotypes", C.c_int32), ("emit_all_genotypes", C.c_int32),
    ("disable_het_snp_rewriting", C.c_int32), ("disable_het_indel_rewriting", C.c_int32),
    ("min_snp_depth", C.c_int32), ("max_snp_depth", C.c_int32), ("min_indel_depth", C.c_int32),
    ("max_indel_depth", C.c_int32),
    ("min_het_snp_quality_by_depth", C.c_float), ("min_hom_snp_quality_by_depth", C.c_float),
    ("min_het_indel_quality_by_depth", C.c_float), ("min_hom_indel_quality_by_depth", C.c_float),
    ("max_snp_phred_strand_bias", C.c_float), ("max_indel_phred_strand_bias", C.c_float),
    ("min_snp_rms_mapping_quality", C.c_float), ("min_indel_rms_mapping_quality", C.c_float),
    ("min_het_snp_alt_allelic_fraction", C.c_float), ("max_het_snp_alt_allelic_fraction", C.c_float),
    ("min_hom_snp_alt_allelic_fraction", C.c_float), ("min_het_indel_alt_allelic_fraction", C.c_float),
    ("max_het_indel_alt_allelic_fraction", C.c_float), ("min_hom_indel_alt_allelic_fraction", C.c_float)]


class FilterParamsStruct(C.Structure):
    _fields_ = FILTER_PARAM_FIELDS


FILTER_OUT_COLUMNS = [("emit", "uint8"), ("rewritten", "uint8"), ("filters_applied", "uint8"),
                      ("filters_failed", "uint32"), ("n_alt", "int32"), ("n_ref", "int32"), ("gq_valid", "uint8")]
FILTER_NAMES = ["HETSNPQD", "HOMSNPQD", "SNPMQ", "SNPFS", "SNPMINDP", "SNPMAXDP", "HETSNPMINAF", "HETSNPMAXAF",
                "HOMSNPMINAF", "HETINDELQD", "HOMINDELQD", "INDELMQ", "INDELFS", "INDELMINDP", "INDELMAXDP",
                "HETINDELMINAF", "HETINDELMAXAF", "HOMINDELMINAF"]


class FilterOutStruct(C.Structure):
    _fields_ = [("n", C.c_int64), ("mem", C.c_int32), ("reserved", C.c_int32)] + \
               [(n, P) for n, _ in FILTER_OUT_COLUMNS]


class SampleRowsStruct(C.Structure):
    _fields_ = [("variants", C.POINTER(SitesStruct)), ("variant_rows", C.POINTER(SiteResultsStruct)),
                ("n_ref_model", C.c_int64), ("ref_model_start", P), ("ref_model_rows", C.POINTER(SiteResultsStruct)),
                ("ref_model_len", P), ("ref_model_max_len", C.c_int32), ("reserved", C.c_int32)]


SQUARE_OUT_COLUMNS = [("kind", "uint8", 1), ("src", "int32", 1), ("n_alt", "int32", 1), ("n_ref", "int32", 1),
                      ("n_other", "int32", 1), ("gl", "float64", "ns")]
SQ_ABSENT, SQ_SCORED, SQ_EXCISED_VARIANT, SQ_EXCISED_REF_MODEL = range(4)


class SquareOutStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_states", C.c_int32), ("mem", C.c_int32)] + \
               [(n, P) for n, _, _ in SQUARE_OUT_COLUMNS]


class TrioStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("mem", C.c_int32), ("reserved", C.c_int32), ("present", P), ("n_alt", P),
                ("n_ref", P), ("n_other", P), ("n_nocall", P), ("kept", P), ("child_action", P), ("filled", P)]


TRIO_UNCHANGED, TRIO_PHASED, TRIO_NO_CALL, TRIO_PHASED_ALT_REF, TRIO_PHASED_REF_ALT = range(5)


class JointInStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_samples", C.c_int32), ("n_states", C.c_int32),
                ("mem", C.c_int32), ("reserved", C.c_int32), ("present", P), ("n_alt", P), ("n_ref", P),
                ("n_other", P), ("n_nocall", P), ("gl", P), ("alt_alleles", P), ("called_alleles", P)]


JOINT_OUT_COLUMNS = [  # name, dtype, shape kind: "N" per site, "SN" per (sample, site), "SNG" x n_states
    ("site_emitted", "uint8", "N"), ("maf", "float64", "N"), ("quality", "float64", "N"),
    ("post0_sum", "float64", "N"), ("flags", "uint8", "SN"), ("posteriors", "float32", "SNG"),
    ("priors", "float32", "SNG"), ("n_alt", "int32", "SN"), ("n_ref", "int32", "SN"), ("gq", "int32", "SN")]
JOINT_POSTERIORS, JOINT_PRIORS, JOINT_RECALLED = 1, 2, 4


class JointOutStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_samples", C.c_int32), ("n_states", C.c_int32),
                ("mem", C.c_int32), ("reserved", C.c_int32)] + [(n, P) for n, _, _ in JOINT_OUT_COLUMNS]


class DepthsInStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_samples", C.c_int32), ("mem", C.c_int32), ("present", P),
                ("read_depth", P), ("ref_read_depth", P), ("sb", P)]


SITE_DEPTH_COLUMNS = ["read_depth", "ref_read_depth", "fwd_read_depth", "rev_read_depth", "ref_fwd_read_depth",
                      "ref_rev_read_depth"]


class SiteDepthsStruct(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("mem", C.c_int32), ("reserved", C.c_int32), ("annotated", P)] + \
               [(n, P) for n in SITE_DEPTH_COLUMNS]


class SynthParamsStruct(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("contig_len", C.c_int32), ("first_pos", C.c_int32),
                ("n_reads_total", C.c_int64), ("read_len", C.c_int32), ("max_indel", C.c_int32),
                ("snp_rate", C.c_double), ("indel_rate", C.c_double), ("triallelic_frac", C.c_double),
                ("softclip_frac", C.c_double), ("sample", C.c_uint64)]


# every symbol declared in include/avocado_b200.h and include/avocado_b200_synth.h
EXPORTED_SYMBOLS = [
    "avo_ctx_create", "avo_ctx_destroy", "avo_last_error", "avo_version", "avo_ctx_set_stream",
    "avo_get_timing", "avo_discover", "avo_call", "avo_discover_and_call",
    "avo_call_all_sites", "avo_filter_genotypes", "avo_square_off", "avo_trio_call", "avo_joint_counts", "avo_joint", "avo_joint_depths",
    "avo_pack_wire", "avo_pack_wire6", "avo_pack_bounds", "avo_pack_reads", "avo_pack_reads_mt", "avo_pack_reads_device", "avo_synth_default_params",
    "avo_synth_host", "avo_synth_device", "avo_synth_text",
    "avo_merge_meta", "avo_shard_table_block_bytes", "avo_shard_rows_block_bytes", "avo_shard_discover", "avo_shard_call", "avo_shard_finish",
    "avo_shard_call_direct", "avo_peer_alloc", "avo_peer_open", "avo_peer_close", "avo_peer_free", "avo_derive_events",
]

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "avocado_b200: %s is missing; there is no CPU fallback. Build it with "
            "`make -C avocado_b200/csrc -j` or `python -c 'import __graft_entry__ as g; g.build()'`."
            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.avo_version.restype = C.c_char_p
    lib.avo_last_error.restype = C.c_char_p
    lib.avo_last_error.argtypes = [P]
    lib.avo_ctx_create.argtypes = [C.c_int, C.POINTER(P)]
    lib.avo_ctx_destroy.argtypes = [P]
    lib.avo_ctx_destroy.restype = None
    lib.avo_ctx_set_stream.argtypes = [P, P]
    lib.avo_get_timing.argtypes = [P, C.POINTER(TimingStruct)]
    lib.avo_discover.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(ParamsStruct),
                                 C.POINTER(SitesOutStruct)]
    lib.avo_call.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(SitesStruct),
                             C.POINTER(CnvStruct), C.POINTER(ParamsStruct),
                             C.POINTER(SiteResultsStruct)]
    lib.avo_discover_and_call.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(CnvStruct),
                                          C.POINTER(ParamsStruct), C.POINTER(SitesOutStruct),
                                          C.POINTER(SiteResultsStruct)]
    lib.avo_call_all_sites.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(SitesStruct),
                                       C.POINTER(CnvStruct), C.POINTER(ParamsStruct),
                                       C.POINTER(SiteResultsStruct), C.POINTER(RefModelOutStruct)]
    lib.avo_filter_genotypes.argtypes = [P, C.POINTER(SiteResultsStruct), P, P, C.POINTER(FilterParamsStruct),
                                         C.POINTER(FilterOutStruct)]
    lib.avo_square_off.argtypes = [P, C.POINTER(SitesStruct), C.POINTER(SampleRowsStruct), C.POINTER(SquareOutStruct)]
    lib.avo_trio_call.argtypes = [P, C.POINTER(TrioStruct)]
    lib.avo_joint_counts.argtypes = [P, C.POINTER(JointInStruct), P, P]
    lib.avo_joint.argtypes = [P, C.POINTER(JointInStruct), C.POINTER(JointOutStruct)]
    lib.avo_joint_depths.argtypes = [P, C.POINTER(DepthsInStruct), C.POINTER(SiteDepthsStruct)]
    lib.avo_pack_wire.argtypes = [C.POINTER(ReadBatchStruct), P, P]
    lib.avo_pack_wire6.argtypes = [C.POINTER(ReadBatchStruct), P, P, P, P, C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32)]
    lib.avo_pack_bounds.argtypes = [C.POINTER(TextReadsStruct)] + [C.POINTER(C.c_int64)] * 4
    lib.avo_pack_reads.argtypes = [C.POINTER(TextReadsStruct), C.POINTER(PackedOutStruct)]
    lib.avo_pack_reads_mt.argtypes = [C.POINTER(TextReadsStruct), C.POINTER(PackedOutStruct), C.c_int]
    lib.avo_pack_reads_device.argtypes = [P, C.POINTER(TextReadsStruct), C.c_int32, C.POINTER(PackedOutStruct)]
    lib.avo_synth_text.argtypes = [C.POINTER(ReadBatchStruct), C.c_int64, C.c_int64, C.POINTER(TextOutStruct)]
    lib.avo_synth_default_params.argtypes = [C.POINTER(SynthParamsStruct)]
    lib.avo_synth_host.argtypes = [C.POINTER(SynthParamsStruct), C.c_int64, C.c_int64,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.POINTER(PackedOutStruct)]
    lib.avo_synth_device.argtypes = [C.POINTER(SynthParamsStruct), C.c_int64, C.c_int64, P, P,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(PackedOutStruct), P]
    lib.avo_merge_meta.argtypes = [P, C.c_int32, P, P, P, P, P, P]
    lib.avo_shard_table_block_bytes.argtypes = [C.POINTER(ShardPlanStruct)]
    lib.avo_shard_table_block_bytes.restype = C.c_int64
    lib.avo_shard_rows_block_bytes.argtypes = [C.POINTER(ShardPlanStruct)]
    lib.avo_shard_rows_block_bytes.restype = C.c_int64
    lib.avo_shard_discover.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(ParamsStruct), C.POINTER(ShardPlanStruct), P]
    lib.avo_shard_call.argtypes = [P, C.POINTER(CnvStruct), C.POINTER(ParamsStruct), C.POINTER(ShardPlanStruct), P,
                                   C.POINTER(SitesOutStruct), P]
    lib.avo_shard_call_direct.argtypes = [P, C.POINTER(CnvStruct), C.POINTER(ParamsStruct), C.POINTER(ShardPlanStruct), P,
                                          C.POINTER(SitesOutStruct), P, C.POINTER(SiteResultsStruct)]
    lib.avo_peer_alloc.argtypes = [P, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.avo_peer_open.argtypes = [P, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.avo_peer_close.argtypes = [P, C.c_void_p]
    lib.avo_peer_free.argtypes = [P, C.c_void_p]
    lib.avo_derive_events.argtypes = [P, C.POINTER(ReadBatchStruct), C.POINTER(ReadEventsStruct)]
    lib.avo_shard_finish.argtypes = [P, C.POINTER(ShardPlanStruct), P, P, C.POINTER(SitesOutStruct),
                                     C.POINTER(SiteResultsStruct), C.POINTER(ShardPlanStruct)]
    _lib = lib
    return lib


class AvocadoError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status
