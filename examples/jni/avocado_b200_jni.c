/* avocado_b200_jni.c -- the JNI glue between org.bdgenomics.avocado.nativelib.AvocadoB200
 * and libavocado_b200.so.  Every Java holder (PackedBatch, Sites, SitesOut, SiteResults, ...) is a plain
 * object whose fields are direct ByteBuffers -- one per column of the struct it mirrors in
 * include/avocado_b200.h -- and long / int scalars; this file only moves addresses and sizes.
 *
 * Build where a JDK exists:  cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *                               examples/jni/avocado_b200_jni.c -Lavocado_b200 -lavocado_b200 -o libavocado_b200_jni.so
 * This image has no JDK: tests/test_packer_cpu.py compiles the file against tests/mock_jni/jni.h (the
 * JDK's names and call shapes) so that it cannot rot. */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "avocado_b200.h"

static void* buf(JNIEnv* e, jobject holder, const char* field) {      /* direct buffer address (NULL if absent) */
  if (!holder) return NULL;
  jclass c = (*e)->GetObjectClass(e, holder);
  jobject b = (*e)->GetObjectField(e, holder, (*e)->GetFieldID(e, c, field, "Ljava/nio/ByteBuffer;"));
  return b ? (*e)->GetDirectBufferAddress(e, b) : NULL;
}
static jlong lng(JNIEnv* e, jobject holder, const char* field) {
  return (*e)->GetLongField(e, holder, (*e)->GetFieldID(e, (*e)->GetObjectClass(e, holder), field, "J"));
}
static jint num(JNIEnv* e, jobject holder, const char* field) {
  return (*e)->GetIntField(e, holder, (*e)->GetFieldID(e, (*e)->GetObjectClass(e, holder), field, "I"));
}
static void set_lng(JNIEnv* e, jobject holder, const char* field, jlong v) {
  (*e)->SetLongField(e, holder, (*e)->GetFieldID(e, (*e)->GetObjectClass(e, holder), field, "J"), v);
}

/* PackedBatch -> avo_read_batch.  mem = AVO_MEM_HOST: the direct buffers were page-locked once with
 * cudaHostRegister; records and operators then cross PCIe in a compact wire form and the payload stays
 * on the host.  The two wire forms are alternatives: the 6-byte records go with THEIR operator columns
 * (wire6 + wire6Oplen + wire6Opcode [+ wire6Exc]), the 16-byte records with the 16-bit operators
 * (wireMeta + wireOps); validate_reads refuses a mix. */
/* ReadEvents holder <-> avo_read_events (the arrays live where the batch lives; nSnv / nIndel: capacities in, counts out) */
static void events_in(JNIEnv* e, jobject je, avo_read_events* v) {
  v->n_snv = lng(e, je, "nSnv"); v->n_indel = lng(e, je, "nIndel");
  v->snv_pos = (int32_t*)buf(e, je, "snvPos"); v->snv_pq = (uint16_t*)buf(e, je, "snvPq");
  v->snv_first = (uint32_t*)buf(e, je, "snvFirst");
  v->indel_read = (int32_t*)buf(e, je, "indelRead"); v->indel_op = (uint32_t*)buf(e, je, "indelOp");
  v->stats.min_start = num(e, je, "minStart"); v->stats.max_end = num(e, je, "maxEnd"); v->stats.max_span = num(e, je, "maxSpan");
  v->stats.valid = num(e, je, "statsValid");
}

/* `ev`: storage for the batch's candidate events when its holder carries a ReadEvents (PackedBatch.events) */
static void read_batch(JNIEnv* e, jobject jr, avo_read_batch* r, avo_read_events* ev) {
  jobject je = (*e)->GetObjectField(e, jr, (*e)->GetFieldID(e, (*e)->GetObjectClass(e, jr), "events",
                                                             "Lorg/bdgenomics/avocado/nativelib/ReadEvents;"));
  if (je && ev) { events_in(e, je, ev); r->events = ev; }
  r->n_reads = lng(e, jr, "nReads"); r->n_blocks = lng(e, jr, "nBlocks");
  r->n_ops = lng(e, jr, "nOps"); r->n_refb = lng(e, jr, "nRefb");
  r->contig_id = num(e, jr, "contigId");
  r->mem = num(e, jr, "mem");
  r->stats.valid = num(e, jr, "statsValid");
  r->stats.min_start = num(e, jr, "minStart"); r->stats.max_end = num(e, jr, "maxEnd"); r->stats.max_span = num(e, jr, "maxSpan");
  r->meta = (const avo_read_meta*)buf(e, jr, "meta");       /* 32-byte records */
  r->payload = (const uint8_t*)buf(e, jr, "payload");       /* 16-byte blocks: 10 bases + 10 qualities */
  r->ops = (const uint32_t*)buf(e, jr, "ops");
  r->refb = (const uint8_t*)buf(e, jr, "refb");
  r->wire6_meta = (const avo_read_wire6*)buf(e, jr, "wire6");
  if (r->wire6_meta) {
    r->wire6_oplen = (const uint8_t*)buf(e, jr, "wire6Oplen");
    r->wire6_opcode = (const uint8_t*)buf(e, jr, "wire6Opcode");
    r->wire6_exc = (const avo_wire_exception*)buf(e, jr, "wire6Exc");
    r->n_wire6_exc = lng(e, jr, "nWire6Exc");
    r->wire6_base = num(e, jr, "wire6Base");
  } else {
    r->wire_meta = (const avo_read_wire*)buf(e, jr, "wireMeta");
    r->wire_ops = (const uint16_t*)buf(e, jr, "wireOps");
  }
}

static void sites_in(JNIEnv* e, jobject js, avo_sites* s) {
  s->n = lng(e, js, "n"); s->n_allele_nibbles = lng(e, js, "nAlleleNibbles"); s->mem = num(e, js, "mem");
  s->contig = (const int32_t*)buf(e, js, "contig"); s->start = (const int32_t*)buf(e, js, "start");
  s->ref_len = (const int32_t*)buf(e, js, "refLen"); s->alt_len = (const int32_t*)buf(e, js, "altLen");
  s->allele_off = (const uint32_t*)buf(e, js, "alleleOff"); s->alleles4 = (const uint8_t*)buf(e, js, "alleles4");
}

static void sites_out(JNIEnv* e, jobject js, avo_sites_out* s) {
  s->capacity = lng(e, js, "capacity"); s->allele_capacity = lng(e, js, "alleleCapacity"); s->mem = num(e, js, "mem");
  s->start = (int32_t*)buf(e, js, "start"); s->ref_len = (int32_t*)buf(e, js, "refLen"); s->alt_len = (int32_t*)buf(e, js, "altLen");
  s->allele_off = (uint32_t*)buf(e, js, "alleleOff"); s->alleles4 = (uint8_t*)buf(e, js, "alleles4");
  s->count = (int32_t*)buf(e, js, "count");
}

static void results(JNIEnv* e, jobject jo, avo_site_results* o) {
  o->n_sites = lng(e, jo, "nSites"); o->n_states = num(e, jo, "nStates"); o->mem = num(e, jo, "mem");
  o->emitted = (uint8_t*)buf(e, jo, "emitted");
  o->allele_fwd = (int32_t*)buf(e, jo, "alleleFwd"); o->other_fwd = (int32_t*)buf(e, jo, "otherFwd");
  o->allele_cov = (int32_t*)buf(e, jo, "alleleCov"); o->other_cov = (int32_t*)buf(e, jo, "otherCov");
  o->total_cov = (int32_t*)buf(e, jo, "totalCov"); o->square_mapq = (double*)buf(e, jo, "squareMapq");
  o->ref_ll = (double*)buf(e, jo, "refLl"); o->allele_ll = (double*)buf(e, jo, "alleleLl");
  o->other_ll = (double*)buf(e, jo, "otherLl"); o->nonref_ll = (double*)buf(e, jo, "nonrefLl");
  o->first_is_ref = (uint8_t*)buf(e, jo, "firstIsRef"); o->copy_number = (int32_t*)buf(e, jo, "copyNumber");
  o->n_alt = (int32_t*)buf(e, jo, "nAlt"); o->n_ref = (int32_t*)buf(e, jo, "nRef"); o->n_other = (int32_t*)buf(e, jo, "nOther");
  o->qual = (double*)buf(e, jo, "qual"); o->gq = (int32_t*)buf(e, jo, "gq"); o->sb = (int32_t*)buf(e, jo, "sb");
  o->fisher = (float*)buf(e, jo, "fisher"); o->rms_mapq = (float*)buf(e, jo, "rmsMapq");
  o->gl = (double*)buf(e, jo, "gl"); o->nonref_gl = (double*)buf(e, jo, "nonrefGl");
}

static void cnv_in(JNIEnv* e, jobject jc, avo_cnv* c) {
  c->n = lng(e, jc, "n");
  c->contig = (const int32_t*)buf(e, jc, "contig"); c->start = (const int32_t*)buf(e, jc, "start");
  c->end = (const int32_t*)buf(e, jc, "end"); c->copy_number = (const int32_t*)buf(e, jc, "copyNumber");
}

static avo_params params(jint ploidy, jint maxQ, jint maxMapQ, jint minPhred, jint minObs) {
  avo_params p = {0};
  p.ploidy = ploidy; p.max_quality = maxQ; p.max_mapq = maxMapQ;
  p.min_phred_discover = minPhred; p.min_obs_discover = minObs;     /* < 0: distinct (no minimum) */
  p.host_payload = 2;      /* page-locked host batches: the call stage's payload blocks are gathered by ... */
  p.gather_threads = 2;    /* ... this many of the library's own host threads */
  return p;
}

JNIEXPORT jlong JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_ctxCreate(JNIEnv* e, jclass k, jint device) {
  avo_ctx* ctx = NULL;
  (void)e; (void)k;
  return avo_ctx_create(device, &ctx) == AVO_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_ctxDestroy(JNIEnv* e, jclass k, jlong ctx) {
  (void)e; (void)k;
  avo_ctx_destroy((avo_ctx*)(intptr_t)ctx);
}

JNIEXPORT jstring JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_lastError(JNIEnv* e, jclass k, jlong ctx) {
  (void)k;
  return (*e)->NewStringUTF(e, avo_last_error((avo_ctx*)(intptr_t)ctx));
}

/* DiscoverVariants.apply (genotyping/DiscoverVariants.scala:51-61) */
JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_discover(
    JNIEnv* e, jclass k, jlong ctx, jobject jr, jint minPhred, jint minObs, jobject js) {
  avo_read_batch r = {0};
  avo_read_events ev = {0};
  avo_sites_out s = {0};
  (void)k;
  read_batch(e, jr, &r, &ev);
  sites_out(e, js, &s);
  avo_params p = params(2, 93, 93, minPhred, minObs);
  const int rc = avo_discover((avo_ctx*)(intptr_t)ctx, &r, &p, &s);
  set_lng(e, js, "n", s.n); set_lng(e, js, "nAlleleNibbles", s.n_allele_nibbles);   /* also the sizes needed on AVO_E_CAPACITY */
  return rc;
}

/* BiallelicGenotyper.call (genotyping/BiallelicGenotyper.scala:88-135), scoreAllSites = false */
JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_call(
    JNIEnv* e, jclass k, jlong ctx, jobject jr, jobject js, jobject jc, jint ploidy, jint maxQ, jint maxMapQ, jobject jo) {
  avo_read_batch r = {0};
  avo_read_events ev = {0};
  avo_sites s = {0};
  avo_cnv c = {0};
  avo_site_results o = {0};
  (void)k;
  read_batch(e, jr, &r, &ev);
  sites_in(e, js, &s);
  if (jc) cnv_in(e, jc, &c);
  results(e, jo, &o);
  avo_params p = params(ploidy, maxQ, maxMapQ, 0, -1);
  return avo_call((avo_ctx*)(intptr_t)ctx, &r, &s, jc ? &c : NULL, &p, &o);
}

/* BiallelicGenotyper.discoverAndCall (BG:156-184) on one region bin */
JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_discoverAndCall(
    JNIEnv* e, jclass k, jlong ctx, jobject jr, jobject jc, jint ploidy, jint maxQ, jint maxMapQ, jint minPhred, jint minObs,
    jobject js, jobject jo) {
  avo_read_batch r = {0};
  avo_read_events ev = {0};
  avo_cnv c = {0};
  avo_sites_out s = {0};
  avo_site_results o = {0};
  (void)k;
  read_batch(e, jr, &r, &ev);
  if (jc) cnv_in(e, jc, &c);
  sites_out(e, js, &s);
  results(e, jo, &o);
  avo_params p = params(ploidy, maxQ, maxMapQ, minPhred, minObs);
  const int rc = avo_discover_and_call((avo_ctx*)(intptr_t)ctx, &r, jc ? &c : NULL, &p, &s, &o);
  set_lng(e, js, "n", s.n); set_lng(e, js, "nAlleleNibbles", s.n_allele_nibbles);
  set_lng(e, jo, "nSites", o.n_sites);
  return rc;
}

/* avo_derive_events: a packer's last step for a batch that will be genotyped more than once */
JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_deriveEvents(JNIEnv* e, jclass k, jlong ctx, jobject jr,
                                                                                     jobject jevents) {
  avo_read_batch r = {0};
  avo_read_events out = {0};
  (void)k;
  read_batch(e, jr, &r, NULL);
  events_in(e, jevents, &out);
  const int rc = avo_derive_events((avo_ctx*)(intptr_t)ctx, &r, &out);
  set_lng(e, jevents, "nSnv", out.n_snv); set_lng(e, jevents, "nIndel", out.n_indel);
  if (rc == AVO_OK) {
    jclass c = (*e)->GetObjectClass(e, jevents);
    (*e)->SetIntField(e, jevents, (*e)->GetFieldID(e, c, "minStart", "I"), out.stats.min_start);
    (*e)->SetIntField(e, jevents, (*e)->GetFieldID(e, c, "maxEnd", "I"), out.stats.max_end);
    (*e)->SetIntField(e, jevents, (*e)->GetFieldID(e, c, "maxSpan", "I"), out.stats.max_span);
    (*e)->SetIntField(e, jevents, (*e)->GetFieldID(e, c, "statsValid", "I"), 1);
  }
  return rc;
}

/* Region-sharded discoverAndCall, one executor per GPU: the three phases; the shim runs ncclAllGather on the
 * two device blocks between them (NCCL's Java binding, or two more native methods around ncclAllGather). */
static void plan_in(JNIEnv* e, jobject jp, avo_shard_plan* p) {
  p->rank = num(e, jp, "rank"); p->world = num(e, jp, "world"); p->own_lo = num(e, jp, "ownLo"); p->own_hi = num(e, jp, "ownHi");
  p->site_cap = lng(e, jp, "siteCap"); p->allele_cap = lng(e, jp, "alleleCap"); p->row_cap = lng(e, jp, "rowCap");
  p->n_states = num(e, jp, "nStates");
}

JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_shardDiscover(
    JNIEnv* e, jclass k, jlong ctx, jobject jr, jint ploidy, jint minPhred, jint minObs, jobject jp, jlong tableBlockDevicePtr) {
  avo_read_batch r = {0};
  avo_read_events ev = {0};
  avo_shard_plan pl = {0};
  (void)k;
  read_batch(e, jr, &r, &ev);
  plan_in(e, jp, &pl);
  avo_params p = params(ploidy, 93, 93, minPhred, minObs);
  return avo_shard_discover((avo_ctx*)(intptr_t)ctx, &r, &p, &pl, (void*)(intptr_t)tableBlockDevicePtr);
}

JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_shardCall(
    JNIEnv* e, jclass k, jlong ctx, jobject jc, jint ploidy, jint minPhred, jint minObs, jobject jp, jlong gatheredTables,
    jobject js, jlong rowsBlockDevicePtr) {
  avo_cnv c = {0};
  avo_shard_plan pl = {0};
  avo_sites_out s = {0};
  (void)k;
  if (jc) cnv_in(e, jc, &c);
  plan_in(e, jp, &pl);
  sites_out(e, js, &s);
  avo_params p = params(ploidy, 93, 93, minPhred, minObs);
  return avo_shard_call((avo_ctx*)(intptr_t)ctx, jc ? &c : NULL, &p, &pl, (const void*)(intptr_t)gatheredTables, &s,
                        (void*)(intptr_t)rowsBlockDevicePtr);
}

/* The same phase with the rows stored straight into the collecting executor's result columns over NVLink
 * (`rootRows`: that executor's own holder there, a holder over the peerOpen mapping elsewhere). */
JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_shardCallDirect(
    JNIEnv* e, jclass k, jlong ctx, jobject jc, jint ploidy, jint minPhred, jint minObs, jobject jp, jlong gatheredTables,
    jobject js, jlong rowsBlockDevicePtr, jobject jrootRows) {
  avo_cnv c = {0};
  avo_shard_plan pl = {0};
  avo_sites_out s = {0};
  avo_site_results root = {0};
  (void)k;
  if (jc) cnv_in(e, jc, &c);
  plan_in(e, jp, &pl);
  sites_out(e, js, &s);
  results(e, jrootRows, &root);
  avo_params p = params(ploidy, 93, 93, minPhred, minObs);
  return avo_shard_call_direct((avo_ctx*)(intptr_t)ctx, jc ? &c : NULL, &p, &pl, (const void*)(intptr_t)gatheredTables, &s,
                               (void*)(intptr_t)rowsBlockDevicePtr, &root);
}

/* Device memory the other executors of the node can map (CUDA IPC): returns the device pointer (0 on failure)
 * and fills the 64-byte handle the shim sends to its peers. */
JNIEXPORT jlong JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_peerAlloc(JNIEnv* e, jclass k, jlong ctx, jlong bytes,
                                                                                   jbyteArray handle) {
  void* ptr = NULL;
  unsigned char h[AVO_PEER_HANDLE_BYTES];
  (void)k;
  if (avo_peer_alloc((avo_ctx*)(intptr_t)ctx, bytes, &ptr, h) != AVO_OK) return 0;
  (*e)->SetByteArrayRegion(e, handle, 0, AVO_PEER_HANDLE_BYTES, (const jbyte*)h);
  return (jlong)(intptr_t)ptr;
}

JNIEXPORT jlong JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_peerOpen(JNIEnv* e, jclass k, jlong ctx, jbyteArray handle) {
  void* ptr = NULL;
  unsigned char h[AVO_PEER_HANDLE_BYTES];
  (void)k;
  (*e)->GetByteArrayRegion(e, handle, 0, AVO_PEER_HANDLE_BYTES, (jbyte*)h);
  return avo_peer_open((avo_ctx*)(intptr_t)ctx, h, &ptr) == AVO_OK ? (jlong)(intptr_t)ptr : 0;
}

JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_peerClose(JNIEnv* e, jclass k, jlong ctx, jlong ptr) {
  (void)e; (void)k;
  return avo_peer_close((avo_ctx*)(intptr_t)ctx, (void*)(intptr_t)ptr);
}

JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_peerFree(JNIEnv* e, jclass k, jlong ctx, jlong ptr) {
  (void)e; (void)k;
  return avo_peer_free((avo_ctx*)(intptr_t)ctx, (void*)(intptr_t)ptr);
}

JNIEXPORT jint JNICALL Java_org_bdgenomics_avocado_nativelib_AvocadoB200_shardFinish(
    JNIEnv* e, jclass k, jlong ctx, jobject jp, jlong gatheredTables, jlong gatheredRows, jobject js, jobject jo, jobject jneeded) {
  avo_shard_plan pl = {0}, need = {0};
  avo_sites_out s = {0};
  avo_site_results o = {0};
  (void)k;
  plan_in(e, jp, &pl);
  sites_out(e, js, &s);
  results(e, jo, &o);
  const int rc = avo_shard_finish((avo_ctx*)(intptr_t)ctx, &pl, (const void*)(intptr_t)gatheredTables,
                                  (const void*)(intptr_t)gatheredRows, &s, &o, &need);
  set_lng(e, js, "n", s.n); set_lng(e, js, "nAlleleNibbles", s.n_allele_nibbles); set_lng(e, jo, "nSites", o.n_sites);
  set_lng(e, jneeded, "siteCap", need.site_cap); set_lng(e, jneeded, "alleleCap", need.allele_cap);
  set_lng(e, jneeded, "rowCap", need.row_cap);
  return rc;      /* AVO_E_CAPACITY / AVO_E_RETRY: every executor runs the three phases again */
}
