// tests/emul/fake_smash.cpp -- TEST INFRASTRUCTURE ONLY.
// A stand-in for libsmash_b200.so, LD_PRELOADed under smash_paper_b200/bin/mummer by the CPU test-suite so that the
// driver's HOST logic (streaming a SAM file / a FASTQ pair through smash_submit_text in chunks: carry-over of
// unconsumed bytes, mate phase, buffer growth, slot rotation, chunk files) can be exercised without a GPU.
// smash_submit_text parses with the host emulation of ingest.cu (emul_ingest.cpp) and "maps" every read to one
// line `name <tab> read_flag <tab> SEQ <tab> QUAL [optional fields]`, which the test compares with the oracle's parse.
// Nothing here is product code; the product has no CPU path.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/smash_b200.h"

extern "C" int emul_ingest(int kind, int final, int replace_n, int phase, const uint8_t *text0, uint64_t n0, const uint8_t *text1, uint64_t n1,
                           uint8_t *names, int64_t *name_off, uint8_t *seq, uint8_t *qual, int64_t *seq_off, uint8_t *opt, int64_t *opt_off,
                           uint16_t *read_flag, uint64_t *out_info, uint64_t *err_index);

struct smash_index { int dummy; };
struct smash_ctx { std::string out[SMASH_N_SLOTS]; uint64_t n_reads[SMASH_N_SLOTS]; bool busy[SMASH_N_SLOTS]; };
static char g_err[256] = "";

extern "C" {
const char *smash_last_error(void) { return g_err; }
int smash_device_count(void) { return 1; }
int smash_index_open(const char *, int, smash_index **out) { *out = new smash_index(); return 0; }
void smash_index_close(smash_index *ix) { delete ix; }
size_t smash_index_sam_header(const smash_index *, char *buf, size_t cap) {
  static const char h[] = "@HD\tfake\n";
  if (buf && cap >= sizeof h - 1) memcpy(buf, h, sizeof h - 1);
  return sizeof h - 1;
}
void smash_params_default(smash_params *p) { memset(p, 0, sizeof *p); p->min_len = 20; p->mode = SMASH_MODE_MAM; }
int smash_ctx_create(const smash_index *, const smash_params *, smash_ctx **out) {
  smash_ctx *c = new smash_ctx();
  for (int s = 0; s < SMASH_N_SLOTS; ++s) { c->busy[s] = false; c->n_reads[s] = 0; }
  *out = c;
  return 0;
}
void smash_ctx_destroy(smash_ctx *c) { delete c; }
void *smash_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void smash_host_free(void *p) { free(p); }

int smash_submit_text(smash_ctx *c, int slot, const smash_text *t, int, smash_text_info *info) {
  if (c->busy[slot]) { snprintf(g_err, sizeof g_err, "slot %d still has a batch in flight", slot); return SMASH_ERR_STATE; }
  const uint64_t n0 = t->n_bytes[0], n1 = t->kind == SMASH_TEXT_FASTQ_PAIR ? t->n_bytes[1] : 0;
  uint64_t lines = 4;
  for (uint64_t i = 0; i < n0; ++i) lines += t->text[0][i] == '\n';
  for (uint64_t i = 0; i < n1; ++i) lines += t->text[1][i] == '\n';
  std::vector<uint8_t> names(n0 + n1 + 64), seq(n0 + n1 + 64), qual(n0 + n1 + 64), opt(n0 + n1 + 64);
  std::vector<int64_t> name_off(lines), seq_off(lines), opt_off(lines);
  std::vector<uint16_t> rf(lines);
  uint64_t inf[8] = {0}, err_index = 0;
  static const uint8_t none = 0;
  const int rc = emul_ingest(t->kind, (t->flags & SMASH_TEXT_FINAL) ? 1 : 0, (t->flags & SMASH_TEXT_REPLACE_N) ? 1 : 0,
                             (t->flags & SMASH_TEXT_MATE2_FIRST) ? 1 : 0, n0 ? (const uint8_t *)t->text[0] : &none, n0,
                             n1 ? (const uint8_t *)t->text[1] : &none, n1, names.data(), name_off.data(), seq.data(), qual.data(),
                             seq_off.data(), opt.data(), opt_off.data(), rf.data(), inf, &err_index);
  if (rc) { snprintf(g_err, sizeof g_err, "input error %d at record %llu", rc, (unsigned long long)err_index); return SMASH_ERR_DATA; }
  std::string &o = c->out[slot];
  o.clear();
  const uint64_t n = inf[0];
  for (uint64_t i = 0; i < n; ++i) {
    o.append((const char *)names.data() + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
    o += '\t'; o += std::to_string(rf[i]); o += '\t';
    o.append((const char *)seq.data() + seq_off[i], (size_t)(seq_off[i + 1] - seq_off[i]));
    o += '\t';
    o.append((const char *)qual.data() + seq_off[i], (size_t)(seq_off[i + 1] - seq_off[i]));
    o.append((const char *)opt.data() + opt_off[i], (size_t)(opt_off[i + 1] - opt_off[i]));
    o += '\n';
  }
  c->n_reads[slot] = n; c->busy[slot] = true;
  if (info) { info->n_reads = n; info->consumed[0] = inf[4]; info->consumed[1] = inf[5]; info->mate2_first_next = (int)inf[6]; }
  return 0;
}
int smash_wait(smash_ctx *c, int slot, smash_result *res) {
  memset(res, 0, sizeof *res);
  res->n_reads = c->n_reads[slot]; res->sam = c->out[slot].data(); res->sam_bytes = c->out[slot].size();
  c->busy[slot] = false;
  return 0;
}
}  // extern "C"
