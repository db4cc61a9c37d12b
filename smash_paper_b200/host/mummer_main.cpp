// mummer_main.cpp -- `mummer`-compatible host driver (C++) over the C ABI of libsmash_b200.so.
//
// Same command line as the reference's mummer.cpp:73-183 (single-dash long flags), same <ref>.bin/
// index files (loaded when present, otherwise built on the GPU and saved in the reference's formats),
// same ./mapout/*.txt output (header of Sequence::sam_header + SAM records), same "Error\n<what>" /
// exit 1 behaviour.  The query reader fills packed batches while the GPU still works on the previous
// one (two slots, smash_submit / smash_wait) -- the counterpart of the reference's reader thread +
// Pair workers (query.cpp:481-520, 614-687).
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "../../include/smash_b200.h"

namespace {

struct Options {
  unsigned min_len = 20;
  int mode = SMASH_MODE_MAM;
  bool nucleotides_only = false, sam_out = false, verbose = false, nomap = false, rcref = false, fastq = false,
       sam_in = false, mappability = false, fastq_pair = false, replace_n = false;
  int threads = 2;
  int gpus = 1;                                   // -gpus N (extension): the query file is cut into N ranges of read pairs
  std::string bins, chromsizes, binout, binstats, mapbin, gc, gcout;   // -bins .. (extension): fused mappability_tag + smashMEM.py + varbin.py
  std::string ref;
  std::vector<std::string> inputs;
};

[[noreturn]] void usage(const char *prog) {
  std::cerr << "Usage: " << prog << " [options] <reference-file> <query-file> ...\n"
            << "  -mum | -mumreference | -mumcand | -maxmatch   match type (default: -mumreference)\n"
            << "  -l N   minimum match length (20)      -n   match only a, c, g, t\n"
            << "  -samin -samout -nomap -rcref -fastq -verbose -qthreads N -mappability\n"
            << "  -minblock N -cached -normalmem (accepted for compatibility)\n"
            << "  -fastqpair [-replaceN] <ref> <mate1.fq[.gz]> <mate2.fq[.gz]>   (extension) the two FASTQ files (plain or gzip) of a mate pair instead of\n"
            << "      `-samin <(fastqs_to_sam mate1.fq mate2.fq [1])`; both are parsed on the GPU\n"
            << "  -gpus N   (extension, with -samin) N GPUs, one contiguous range of read pairs each\n"
            << "  -bins <bins.txt> -chromsizes <chrom_sizes.txt> -binout <varbin.txt> [-binstats <stats.txt>] [-mapbin <map.bin>]\n"
            << "      (extension) the stages after mummer in smash_mapping.sh / binning.sh on the GPU(s): mappability_tag,\n"
            << "      smashMEM.py <bam> 0 0 10000 4, the chromosome filter and varbin.py; counts of all GPUs are summed by one\n"
            << "      NCCL allreduce\n"
            << "  -gc <gc.txt> -gcout <lowratio.txt>   (extension, with -bins) GC normalisation of the counts, the head of cbs.r\n"
            << "      (cbs.r:18-25): rows `chrom chrompos abspos bincount ratio gc.content lowratio`\n";
  std::exit(1);
}

Options parse(int argc, char **argv) {
  Options o;
  enum { L = 1, MUMREF, MAXMATCH, MUM, MUMCAND, N, QTHREADS, SAMOUT, VERBOSE, NOMAP, RCREF, FASTQ, SAMIN, MAPPABILITY, CACHED, NORMALMEM, MINBLOCK, FASTQPAIR, REPLACEN, GPUS, BINS, CHROMSIZES, BINOUT, BINSTATS, MAPBIN, GC, GCOUT };
  static const option table[] = {
      {"l", required_argument, nullptr, L},          {"mumreference", no_argument, nullptr, MUMREF},
      {"maxmatch", no_argument, nullptr, MAXMATCH},  {"mum", no_argument, nullptr, MUM},
      {"mumcand", no_argument, nullptr, MUMCAND},    {"n", no_argument, nullptr, N},
      {"qthreads", required_argument, nullptr, QTHREADS}, {"samout", no_argument, nullptr, SAMOUT},
      {"verbose", no_argument, nullptr, VERBOSE},    {"nomap", no_argument, nullptr, NOMAP},
      {"rcref", no_argument, nullptr, RCREF},        {"fastq", no_argument, nullptr, FASTQ},
      {"samin", no_argument, nullptr, SAMIN},        {"mappability", no_argument, nullptr, MAPPABILITY},
      {"cached", no_argument, nullptr, CACHED},      {"normalmem", no_argument, nullptr, NORMALMEM},
      {"minblock", required_argument, nullptr, MINBLOCK}, {"fastqpair", no_argument, nullptr, FASTQPAIR},
      {"replaceN", no_argument, nullptr, REPLACEN},  {"gpus", required_argument, nullptr, GPUS},
      {"bins", required_argument, nullptr, BINS},    {"chromsizes", required_argument, nullptr, CHROMSIZES},
      {"binout", required_argument, nullptr, BINOUT}, {"binstats", required_argument, nullptr, BINSTATS},
      {"mapbin", required_argument, nullptr, MAPBIN}, {"gc", required_argument, nullptr, GC},
      {"gcout", required_argument, nullptr, GCOUT}, {nullptr, 0, nullptr, 0}};
  for (;;) {
    int idx = -1;
    const int c = getopt_long_only(argc, argv, "", table, &idx);
    if (c == -1) break;
    switch (c) {
      case L: o.min_len = (unsigned)atol(optarg); break;
      case MUMREF: case MUMCAND: o.mode = SMASH_MODE_MAM; break;
      case MAXMATCH: o.mode = SMASH_MODE_MEM; break;
      case MUM: o.mode = SMASH_MODE_MUM; break;
      case N: o.nucleotides_only = true; break;
      case QTHREADS: o.threads = atoi(optarg); break;
      case SAMOUT: o.sam_out = true; break;
      case VERBOSE: o.verbose = true; break;
      case NOMAP: o.nomap = true; break;
      case RCREF: o.rcref = true; break;
      case FASTQ: o.fastq = true; break;
      case FASTQPAIR: o.fastq_pair = true; break;
      case REPLACEN: o.replace_n = true; break;
      case SAMIN: o.sam_in = true; break;
      case MAPPABILITY: o.mappability = true; break;
      case CACHED: case NORMALMEM: case MINBLOCK: break;
      case GPUS: o.gpus = atoi(optarg); break;
      case BINS: o.bins = optarg; break;
      case CHROMSIZES: o.chromsizes = optarg; break;
      case BINOUT: o.binout = optarg; break;
      case BINSTATS: o.binstats = optarg; break;
      case MAPBIN: o.mapbin = optarg; break;
      case GC: o.gc = optarg; break;
      case GCOUT: o.gcout = optarg; break;
      default: std::cerr << "Invalid arguments." << std::endl; usage(argv[0]);
    }
  }
  if (argc - optind < 2) { std::cerr << "There are too few arguments" << std::endl; usage(argv[0]); }
  if (o.fastq && o.sam_in) throw std::runtime_error("-fastq cannot be used with -samin");
  if (o.fastq_pair && (o.fastq || o.sam_in)) throw std::runtime_error("-fastqpair cannot be used with -fastq or -samin");
  if (o.fastq_pair && argc - optind != 3) throw std::runtime_error("-fastqpair takes the reference and exactly two fastq files");
  if (o.replace_n && !o.fastq_pair) throw std::runtime_error("-replaceN can only be used with -fastqpair");
  if (o.nomap && !o.sam_out) throw std::runtime_error("-nomap can only be used with -sam_out");
  if (o.mappability && !o.rcref) throw std::runtime_error("-mappability requires -rcref");
  if (o.gpus < 1) throw std::runtime_error("-gpus must be at least 1");
  if (o.gpus > 1 && !o.sam_in) throw std::runtime_error("-gpus needs -samin (the query file is cut into ranges of SAM lines)");
  if (!o.bins.empty() && (o.chromsizes.empty() || o.binout.empty())) throw std::runtime_error("-bins needs -chromsizes and -binout");
  if (o.gc.empty() != o.gcout.empty() || (!o.gc.empty() && o.bins.empty())) throw std::runtime_error("-gc and -gcout go together and need -bins");
  if (!o.bins.empty() && (!o.rcref || o.mode != SMASH_MODE_MAM)) throw std::runtime_error("-bins needs -rcref and the default match type");
  o.ref = argv[optind];
  for (int i = optind + 1; i < argc; ++i) o.inputs.push_back(argv[i]);
  return o;
}

void check(int rc) { if (rc) throw std::runtime_error(smash_last_error()); }

// ---- reference text (Sequence build branch, fasta.cpp:138-203) ------------------------------------
char complement(char c) {
  switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'r': return 'y'; case 'y': return 'r'; case 'm': return 'k'; case 'k': return 'm';
    case 'b': return 'v'; case 'd': return 'h'; case 'h': return 'd'; case 'v': return 'b';
    default: return c;
  }
}
struct RefText {
  std::string text; std::vector<uint64_t> startpos, sizes; std::vector<std::string> descr;
};
RefText read_fasta(const std::string &path, bool rcref, bool verbose) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("unable to open " + path);
  std::vector<std::string> names, seqs;
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    size_t b = 0; while (b < line.size() && line[b] == ' ') ++b;
    if (b == line.size()) continue;
    if (line[b] == '>') {
      size_t s = b + 1; while (s < line.size() && line[s] == ' ') ++s;
      size_t e = line.find(' ', s);
      names.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s));
      seqs.emplace_back();
    } else if (!seqs.empty()) {
      for (size_t i = b; i < line.size(); ++i) seqs.back().push_back((char)tolower((unsigned char)line[i]));
    }
  }
  RefText r;
  for (size_t k = 0; k < names.size(); ++k) {
    const bool last = k + 1 == names.size();
    const std::string &s = seqs[k];
    if (verbose) std::cerr << "# " << names[k] << " " << s.size() << " " << r.text.size() << std::endl;
    r.startpos.push_back(r.text.size()); r.sizes.push_back(s.size()); r.descr.push_back(names[k]);
    r.text += s;
    if (rcref || !last) r.text.push_back('`');
    if (rcref) {
      r.startpos.push_back(r.text.size()); r.sizes.push_back(s.size()); r.descr.push_back(names[k]);
      for (size_t i = s.size(); i-- > 0;) r.text.push_back(complement(s[i]));
      if (!last) r.text.push_back('`');
    }
  }
  r.text.push_back('$');
  return r;
}

// ---- packed batch (what crosses the ABI) ------------------------------------------------------------
struct Batch {
  std::vector<uint8_t> names, seq, qual, opt;
  std::vector<int64_t> name_off{0}, seq_off{0}, opt_off{0};
  std::vector<uint16_t> read_flag;
  size_t n() const { return read_flag.size(); }
  void clear() { names.clear(); seq.clear(); qual.clear(); opt.clear(); name_off.assign(1, 0); seq_off.assign(1, 0); opt_off.assign(1, 0); read_flag.clear(); }
  // QueryReader::run + Aligner::reset (query.cpp:643-644, 185-201)
  void add(std::string name, unsigned flag, bool from_flag, const std::string &s, const std::string &q, const std::string &optional) {
    if (from_flag) { if (flag & 64) name += ":0"; else if (flag & 128) name += ":1"; }
    uint16_t rf = 0;
    if (name.size() >= 2 && name[name.size() - 2] == ':') {
      if (name.back() == '0') { name.resize(name.size() - 2); rf = 65; }
      else if (name.back() == '1') { name.resize(name.size() - 2); rf = 129; }
    }
    names.insert(names.end(), name.begin(), name.end()); name_off.push_back((int64_t)names.size());
    size_t end = s.size(); while (end && s[end - 1] == ' ') --end;
    size_t len = 0;
    for (size_t i = 0; i < end; ++i) if (s[i] != ' ') { seq.push_back((uint8_t)s[i]); ++len; }
    std::string qq = q.empty() ? std::string(len, '!') : q;          // Aligner::run, query.cpp:323
    qq.resize(len, '!');
    qual.insert(qual.end(), qq.begin(), qq.end());
    seq_off.push_back((int64_t)seq.size());
    opt.insert(opt.end(), optional.begin(), optional.end()); opt_off.push_back((int64_t)opt.size());
    read_flag.push_back(rf);
  }
  smash_batch view(uint64_t first_pair) const {
    smash_batch b{};
    b.n_reads = n(); b.names = names.data(); b.name_off = name_off.data(); b.seq = seq.data(); b.qual = qual.data();
    b.seq_off = seq_off.data(); b.opt = opt.empty() ? nullptr : opt.data(); b.opt_off = opt.empty() ? nullptr : opt_off.data();
    b.read_flag = read_flag.data(); b.first_pair_ordinal = first_pair;
    return b;
  }
};

// one input record per call; false at end of file
struct QueryParser {
  std::ifstream in; const Options &o;
  QueryParser(const std::string &path, const Options &opt) : in(path), o(opt) { if (!in) throw std::runtime_error("unable to open " + path); }
  bool next(Batch &b) {
    std::string line;
    while (std::getline(in, line)) {
      if (line.empty()) continue;
      if (o.sam_in) {                                    // query.cpp:639-648
        std::istringstream f(line);
        std::string name, ref, pos, mapq, cigar, mref, mpos, tlen, seq, errors, tok, optional;
        unsigned flag = 0;
        f >> name >> flag >> ref >> pos >> mapq >> cigar >> mref >> mpos >> tlen >> seq >> errors;
        while (f >> tok) { optional += "\t"; optional += tok; }
        b.add(name, flag, true, seq, errors, optional);
        return true;
      }
      const char start = o.fastq ? '@' : '>';          // query.cpp:649-680
      if (line[0] != start) throw std::runtime_error(std::string("missing query start character ") + start + " in input line " + line);
      size_t s = 1; while (s < line.size() && line[s] == ' ') ++s;
      size_t e = line.size(); while (e > s && line[e - 1] == ' ') --e;
      std::string meta;
      for (size_t i = s; i < e; ++i) {
        if (line[i] == ' ') { if (i + 1 != e) { if (line[i + 1] == '1') meta += ":0"; else if (line[i + 1] == '2') meta += ":1"; } break; }
        meta += line[i];
      }
      std::string seq, errors, plus;
      if (!std::getline(in, seq) || seq.empty()) throw std::runtime_error("empty sequence");
      if (o.fastq) { std::getline(in, plus); std::getline(in, errors); if (errors.empty()) throw std::runtime_error("empty errors"); }
      b.add(meta, 0, false, seq, errors, "");
      return true;
    }
    return false;
  }
};

// first line start at or after `approx` whose SAM flag marks a first mate (or an unpaired read): a cut there keeps
// reads 2k and 2k+1 of every pair in one range (query.cpp:629-637 pairs by arrival)
uint64_t pair_boundary(FILE *f, uint64_t approx, uint64_t size) {
  if (approx == 0) return 0;
  fseeko(f, (off_t)(approx - 1), SEEK_SET);
  uint64_t pos = approx - 1;
  int ch;
  while ((ch = fgetc(f)) != EOF) { ++pos; if (ch == '\n') break; }          // to the start of the next line
  std::string line;
  for (;;) {
    if (pos >= size) return size;
    line.clear();
    while ((ch = fgetc(f)) != EOF && ch != '\n') line.push_back((char)ch);
    const uint64_t next = pos + line.size() + (ch == '\n' ? 1 : 0);
    const size_t t1 = line.find('\t');
    const unsigned flag = t1 == std::string::npos ? 0u : (unsigned)strtoul(line.c_str() + t1 + 1, nullptr, 10);
    if (!line.empty() && line[0] != '@' && !(flag & 128)) return pos;        // first mate (64) or unpaired
    if (ch == EOF) return size;
    pos = next;
  }
}
// str(float) of Python 3 (varbin.py:97 writes the ratio column with it): shortest digits that round-trip, fixed
// notation for 1e-4 <= |x| < 1e16, always a fractional part
std::string py_repr(double v) {
  if (v != v) return "nan";
  if (v == 1.0 / 0.0) return "inf";
  if (v == -1.0 / 0.0) return "-inf";
  char buf[64];
  int p = 1;
  for (; p <= 17; ++p) { snprintf(buf, sizeof buf, "%.*e", p - 1, v); if (strtod(buf, nullptr) == v) break; }
  std::string m(buf);
  const size_t epos = m.find('e');
  const int e10 = atoi(m.c_str() + epos + 1);
  std::string digits; bool neg = false;
  for (size_t i = 0; i < epos; ++i) { if (m[i] == '-') neg = true; else if (m[i] != '.') digits.push_back(m[i]); }
  std::string out = neg ? "-" : "";
  if (e10 >= -4 && e10 < 16) {
    if (e10 >= 0) {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)e10 + 1));
      ip.append((size_t)e10 + 1 - ip.size(), '0');
      std::string fp = digits.size() > (size_t)e10 + 1 ? digits.substr((size_t)e10 + 1) : "0";
      out += ip + "." + fp;
    } else out += "0." + std::string((size_t)(-e10 - 1), '0') + digits;
  } else {
    out += digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    char eb[16]; snprintf(eb, sizeof eb, "e%c%02d", e10 < 0 ? '-' : '+', e10 < 0 ? -e10 : e10);
    out += eb;
  }
  return out;
}
// varbin.py:95-114: rows `chrom start abs count ratio`, then the stats file
void write_varbin(const Options &o, const std::vector<std::string> &rows, std::vector<int64_t> counts, const smash_tail_stats &st) {
  FILE *f = fopen(o.binout.c_str(), "wb");
  if (!f) throw std::runtime_error("could not open " + o.binout + " for writing");
  const double per_bin = (double)st.reads_kept / (double)rows.size();
  for (size_t i = 0; i < rows.size(); ++i) {
    const std::string r = per_bin > 0 ? py_repr((double)counts[i] / per_bin) : std::string("nan");
    fprintf(f, "%s\t%lld\t%s\n", rows[i].c_str(), (long long)counts[i], r.c_str());
  }
  fclose(f);
  if (!o.binstats.empty()) {
    FILE *s = fopen(o.binstats.c_str(), "wb");
    if (!s) throw std::runtime_error("could not open " + o.binstats + " for writing");
    std::sort(counts.begin(), counts.end());
    fprintf(s, "TotalReads\tDupsRemoved\tReadsKept\tMedianBinCount\n%llu\t%llu\t%llu\t%lld\n", (unsigned long long)st.total_reads,
            (unsigned long long)st.dups_removed, (unsigned long long)st.reads_kept, (long long)(counts.empty() ? 0 : counts[counts.size() / 2]));
    fclose(s);
  }
}

// cbs.r:11-25: gc.txt (header line; columns bin.chrom and gc.content) + the counts -> ratio and lowratio per bin, computed on
// the GPU (smash_gcnorm_*).  One row per bin: varbin's three leading columns, the count, then the three doubles.
void write_gcnorm(const Options &o, const std::vector<std::string> &rows, const std::vector<int64_t> &counts) {
  std::ifstream in(o.gc);
  if (!in) throw std::runtime_error("unable to open " + o.gc);
  std::string line;
  if (!std::getline(in, line)) throw std::runtime_error(o.gc + ": empty file");
  int c_chrom = -1, c_gc = -1, col = 0;
  { std::istringstream hs(line); std::string t; while (hs >> t) { if (t == "bin.chrom") c_chrom = col; if (t == "gc.content") c_gc = col; ++col; } }
  if (c_chrom < 0 || c_gc < 0) throw std::runtime_error(o.gc + ": header needs the columns bin.chrom and gc.content (cbs.r:11-12, 23)");
  std::vector<double> gc; std::vector<uint8_t> autosome;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    std::istringstream ls(line); std::string t, chrom, g; col = 0;
    while (ls >> t) { if (col == c_chrom) chrom = t; if (col == c_gc) g = t; ++col; }
    if (chrom.empty() || g.empty()) throw std::runtime_error(o.gc + ": short line");
    gc.push_back(strtod(g.c_str(), nullptr));
    // chrom.numeric < 23 (cbs.r:13-16, 21): chr1..chr22
    const std::string num = chrom.size() > 3 ? chrom.substr(3) : std::string();
    char *end = nullptr; const long v = strtol(num.c_str(), &end, 10);
    autosome.push_back((!num.empty() && *end == 0 && v < 23) ? 1 : 0);
  }
  if (gc.size() != counts.size()) throw std::runtime_error(o.gc + ": " + std::to_string(gc.size()) + " bins, the bin file has " + std::to_string(counts.size()));
  smash_gcnorm *g = nullptr;
  check(smash_gcnorm_create(0, gc.data(), autosome.data(), gc.size(), 0.05, 3, &g));
  std::vector<double> ratio(gc.size()), low(gc.size());
  const int rc = smash_gcnorm_run(g, counts.data(), nullptr, ratio.data(), low.data());
  smash_gcnorm_destroy(g);
  check(rc);
  FILE *f = fopen(o.gcout.c_str(), "wb");
  if (!f) throw std::runtime_error("could not open " + o.gcout + " for writing");
  fprintf(f, "chrom\tchrompos\tabspos\tbincount\tratio\tgc.content\tlowratio\n");
  for (size_t i = 0; i < gc.size(); ++i)
    fprintf(f, "%s\t%lld\t%.15g\t%.15g\t%.15g\n", rows[i].c_str(), (long long)counts[i], ratio[i], gc[i], low[i]);
  fclose(f);
}

bool readable(const std::string &p) { return access(p.c_str(), R_OK) == 0; }

}  // namespace

int main(int argc, char **argv) {
  try {
    const Options o = parse(argc, argv);
    smash_params p; smash_params_default(&p);
    p.mode = o.mode; p.min_len = o.min_len; p.nomap = o.nomap; p.nucleotides_only = o.nucleotides_only;
    const std::string base = o.ref + ".bin/rc" + (o.rcref ? "1" : "0");
    const bool have_index = readable(base + ".i4.index.bin") || readable(base + ".i8.index.bin");
    smash_index *ix = nullptr; smash_ctx *ctx = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (have_index) {
      if (o.verbose) std::cerr << "# loading reference binary\n# loading index binary" << std::endl;
      check(smash_index_open(o.ref.c_str(), o.rcref, &ix));
      check(smash_ctx_create(ix, &p, &ctx));
    } else {
      if (o.verbose) std::cerr << "# loading reference from fasta" << std::endl;
      RefText r = read_fasta(o.ref, o.rcref, o.verbose);
      if (o.verbose) std::cerr << "# seq_vec.length=" << r.text.size() << "\n# creating index from reference" << std::endl;
      std::vector<const char *> d; for (auto &s : r.descr) d.push_back(s.c_str());
      const uint64_t N = r.text.size();
      const int w = N >= 0xffffffffull - 100000 ? 8 : 4;             // mummer.cpp:156-183
      check(smash_ctx_create_from_text((const uint8_t *)r.text.data(), N, r.descr.size(), r.startpos.data(), r.sizes.data(),
                                       d.data(), o.rcref, w, 1, 0, &p, &ctx));
      if (o.verbose) std::cerr << "# saving index" << std::endl;
      check(smash_ctx_save_index(ctx, o.ref.c_str(), 0));
      check(smash_index_open(o.ref.c_str(), o.rcref, &ix));
    }
    if (o.verbose)
      std::cerr << "# constructed index in "
                << std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count() << " seconds" << std::endl;
    if (o.mappability) {                                             // mummer.cpp:50-53
      check(smash_ctx_build_mappability(ctx, nullptr, 0));
      // total forward bases = sum of @SQ lengths
      std::vector<char> hdr(smash_index_sam_header(ix, nullptr, 0));
      smash_index_sam_header(ix, hdr.data(), hdr.size());
      uint64_t total = 0;
      { std::istringstream hs(std::string(hdr.begin(), hdr.end())); std::string l;
        while (std::getline(hs, l)) { size_t k = l.find("\tLN:"); if (l.rfind("@SQ", 0) == 0 && k != std::string::npos) total += strtoull(l.c_str() + k + 4, nullptr, 10); } }
      std::vector<uint8_t> body(2 * total);
      check(smash_ctx_build_mappability(ctx, body.data(), body.size()));
      FILE *f = fopen(o.inputs[0].c_str(), "wb");
      if (!f) throw std::runtime_error("could not open outfile " + o.inputs[0] + " for writing");
      fputc(0, f); fputc(0, f);                                      // the reference's two junk bytes
      fwrite(body.data(), 1, body.size(), f); fclose(f);
      smash_ctx_destroy(ctx); smash_index_close(ix);
      return 0;
    }
    // -gpus N: one context per GPU over the same host index; -bins: map.bin + the tail's configuration on each
    const int n_gpus = o.gpus;
    std::vector<smash_ctx *> ctxs{ctx};
    for (int g = 1; g < n_gpus; ++g) { smash_params pg = p; pg.device = g; smash_ctx *cg = nullptr; check(smash_ctx_create(ix, &pg, &cg)); ctxs.push_back(cg); }
    std::vector<int64_t> bin_starts; std::vector<std::string> bin_rows;
    if (!o.bins.empty()) {
      std::ifstream bf(o.bins); if (!bf) throw std::runtime_error("unable to open " + o.bins);
      std::string l;
      while (std::getline(bf, l)) {                                  // bins.txt: chrom, start, abs start, ... (varbin.py:17-26)
        if (l.empty()) continue;
        std::istringstream ls(l); std::string c0, c1, c2; ls >> c0 >> c1 >> c2;
        bin_rows.push_back(c0 + "\t" + c1 + "\t" + c2); bin_starts.push_back(strtoll(c2.c_str(), nullptr, 10));
      }
      std::ifstream cf(o.chromsizes); if (!cf) throw std::runtime_error("unable to open " + o.chromsizes);
      std::vector<std::string> cnames; std::vector<int64_t> coffs;
      while (std::getline(cf, l)) {                                  // chrom_sizes.txt: name, size, cumulative offset (varbin.py:28-36)
        if (l.empty()) continue;
        std::istringstream ls(l); std::string c0, c1, c2; ls >> c0 >> c1 >> c2;
        cnames.push_back(c0); coffs.push_back(strtoll(c2.c_str(), nullptr, 10));
      }
      std::vector<const char *> cptr; for (auto &c : cnames) cptr.push_back(c.c_str());
      const std::string mb = o.mapbin.empty() ? o.ref + ".bin/map.bin" : o.mapbin;
      std::ifstream mf(mb, std::ios::binary); if (!mf) throw std::runtime_error("unable to open " + mb + " (mummer -mappability writes it)");
      std::vector<char> body((std::istreambuf_iterator<char>(mf)), std::istreambuf_iterator<char>());
      if (body.size() < 2) throw std::runtime_error(mb + " is empty");
      for (smash_ctx *cg : ctxs) {
        check(smash_ctx_load_mappability(cg, (const uint8_t *)body.data() + 2, body.size() - 2));   // 2 junk bytes (longSA.cpp:606-617)
        check(smash_tail_configure(cg, bin_starts.data(), bin_starts.size(), cptr.data(), coffs.data(), cnames.size(), 10000, 4));
      }
    }
    if (n_gpus > 1) check(smash_comm_init_all(ctxs.data(), n_gpus));
    std::vector<char> header(smash_index_sam_header(ix, nullptr, 0));
    smash_index_sam_header(ix, header.data(), header.size());
    if (o.verbose) std::cerr << "# running " << o.threads << " threads to answer queries\n# running " << o.inputs.size() << " query reader" << std::endl;
    const size_t BATCH = 1u << 20;                                   // reads per batch (even)
    // every batch becomes one chunk file, its lines sorted as OutputSorter::flush sorts a chunk (query.cpp:448-468);
    // SMASH_CHUNK_ORDER=input keeps them in input order instead
    const char *ord_env = getenv("SMASH_CHUNK_ORDER");
    const int want_sam = SMASH_WANT_SAM | ((ord_env && !strcmp(ord_env, "input")) ? 0 : SMASH_WANT_SORTED) | (o.bins.empty() ? 0 : SMASH_WANT_TAIL);
    const auto tq = std::chrono::steady_clock::now();
    const bool device_reader = o.sam_in && !getenv("SMASH_HOST_READER");   // -samin text is parsed on the GPU (smash_submit_text)
    if (n_gpus > 1 && !device_reader) throw std::runtime_error("-gpus needs the device reader (-samin without SMASH_HOST_READER)");
    std::atomic<uint64_t> total_queries{0};
    // one GPU's share of the work: with -gpus N, GPU g takes the g-th of N byte ranges of every query file, cut at
    // read-pair boundaries (contiguous ranges of pairs: what the sharded tail needs)
    auto run_gpu = [&](int g) {
    smash_ctx *ctx = ctxs[g];
    uint64_t n_queries = 0, chunk = 0, pairs_done = 0;
    for (size_t fi = 0; fi < o.inputs.size(); ++fi) {
      if (o.fastq_pair && fi > 0) break;                            // the two inputs are ONE stream of pairs
      bool in_flight[SMASH_N_SLOTS] = {};
      auto drain = [&](int slot) {
        smash_result r; check(smash_wait(ctx, slot, &r));
        in_flight[slot] = false;
        if (o.sam_out && r.sam_bytes) {
          mkdir("mapout", 0755);
          const std::string fn = "mapout/mapoutb200_" + (n_gpus > 1 ? "g" + std::to_string(g) + "_" : std::string()) + std::to_string(fi) + "." + std::to_string(++chunk) + ".txt";
          FILE *f = fopen(fn.c_str(), "wb");
          if (!f) throw std::runtime_error("Problem opening out file");
          fwrite(header.data(), 1, header.size(), f); fwrite(r.sam, 1, r.sam_bytes, f); fclose(f);
        }
      };
      int slot = 0; bool more = true;
      if (o.fastq_pair) {
        // fastqs_to_sam's loop (fastqs_to_sam.cpp:47-95) + the SAM reader, both on the device: the two files are
        // streamed through two pinned buffers; each call says how much of either it used and which mate the next
        // chunk starts with.
        // The files may be gzip streams (smash_mapping.sh:19 feeds fastqs_to_sam from two `zcat`s): zlib inflates them
        // here, one thread per mate file as the two zcat processes of the script run side by side; plain text passes
        // through gzread unchanged.
        gzFile f[2]; char *text[2]; size_t cap[2], len[2] = {0, 0}; bool eof[2] = {false, false};
        size_t chunk_bytes = (size_t)128 << 20;
        if (const char *e = getenv("SMASH_TEXT_CHUNK")) chunk_bytes = std::max<size_t>(4096, strtoull(e, nullptr, 10));
        for (int k = 0; k < 2; ++k) {
          f[k] = gzopen(o.inputs[k].c_str(), "rb");
          if (!f[k]) throw std::runtime_error("Could not open fastq file " + o.inputs[k]);      // fastqs_to_sam.cpp:35-37
          gzbuffer(f[k], 1u << 20);
          cap[k] = chunk_bytes; text[k] = (char *)smash_host_alloc(cap[k]);
          if (!text[k]) throw std::runtime_error("out of pinned host memory");
        }
        int mate2_first = 0;
        for (;;) {
          {
            bool bad[2] = {false, false};
            auto fill = [&](int k) {
              while (len[k] < cap[k] && !eof[k]) {
                const int got = gzread(f[k], text[k] + len[k], (unsigned)std::min<size_t>(cap[k] - len[k], (size_t)1 << 30));
                if (got < 0) { bad[k] = true; return; }
                len[k] += (size_t)got; if (!got) eof[k] = true;
              }
            };
            std::thread other(fill, 1);
            fill(0);
            other.join();
            for (int k = 0; k < 2; ++k) if (bad[k]) throw std::runtime_error("error reading fastq file " + o.inputs[k] + " (corrupt gzip stream?)");
          }
          const bool final = eof[0] && eof[1];
          if (in_flight[slot]) drain(slot);
          smash_text t{}; smash_text_info info{};
          t.kind = SMASH_TEXT_FASTQ_PAIR;
          t.flags = (final ? SMASH_TEXT_FINAL : 0) | (o.replace_n ? SMASH_TEXT_REPLACE_N : 0) | (mate2_first ? SMASH_TEXT_MATE2_FIRST : 0);
          for (int k = 0; k < 2; ++k) { t.text[k] = text[k]; t.n_bytes[k] = len[k]; }
          t.first_pair_ordinal = pairs_done;
          check(smash_submit_text(ctx, slot, &t, want_sam, &info));
          in_flight[slot] = true; n_queries += info.n_reads; pairs_done += (info.n_reads + 1) / 2;
          mate2_first = info.mate2_first_next;
          slot = (slot + 1) % SMASH_N_SLOTS;
          if (final) break;
          for (int k = 0; k < 2; ++k) {
            if (!eof[k] && info.consumed[k] == 0 && len[k] == cap[k]) {     // nothing of a full buffer was usable: grow it
              char *bigger = (char *)smash_host_alloc(2 * cap[k]);
              if (!bigger) throw std::runtime_error("out of pinned host memory");
              memcpy(bigger, text[k], len[k]); smash_host_free(text[k]); text[k] = bigger; cap[k] *= 2;
            }
            len[k] -= info.consumed[k];
            memmove(text[k], text[k] + info.consumed[k], len[k]);
          }
        }
        for (int k = 0; k < 2; ++k) { gzclose(f[k]); smash_host_free(text[k]); }
        more = false;
      } else if (device_reader) {
        // QueryReader::run's SAM branch (query.cpp:625-648) on the device: the file is streamed through a pinned
        // buffer in large chunks; what a chunk leaves unconsumed (a cut line, an odd trailing read) is carried over.
        FILE *f = fopen(o.inputs[fi].c_str(), "rb");
        if (!f) throw std::runtime_error("unable to open " + o.inputs[fi]);
        uint64_t range_begin = 0, range_end = UINT64_MAX;
        if (n_gpus > 1) {
          struct stat fst; if (stat(o.inputs[fi].c_str(), &fst)) throw std::runtime_error("unable to open " + o.inputs[fi]);
          range_begin = pair_boundary(f, (uint64_t)fst.st_size / n_gpus * g, (uint64_t)fst.st_size);
          range_end = g + 1 == n_gpus ? (uint64_t)fst.st_size : pair_boundary(f, (uint64_t)fst.st_size / n_gpus * (g + 1), (uint64_t)fst.st_size);
          fseeko(f, (off_t)range_begin, SEEK_SET);
        }
        uint64_t left_in_range = range_end - range_begin;
        size_t cap = (size_t)256 << 20, len = 0;
        if (const char *e = getenv("SMASH_TEXT_CHUNK")) cap = std::max<size_t>(4096, strtoull(e, nullptr, 10));
        char *text = (char *)smash_host_alloc(cap);
        if (!text) throw std::runtime_error("out of pinned host memory");
        bool eof = false;
        while (!eof) {
          while (len < cap && !eof) {
            const size_t ask = (uint64_t)(cap - len) < left_in_range ? cap - len : (size_t)left_in_range;
            const size_t got = ask ? fread(text + len, 1, ask, f) : 0;
            len += got; left_in_range -= got;
            if (!got) eof = true;
          }
          if (in_flight[slot]) drain(slot);
          smash_text t{}; smash_text_info info{};
          t.kind = SMASH_TEXT_SAM; t.flags = eof ? SMASH_TEXT_FINAL : 0; t.text[0] = text; t.n_bytes[0] = len;
          t.first_pair_ordinal = pairs_done;
          check(smash_submit_text(ctx, slot, &t, want_sam, &info));
          in_flight[slot] = true; n_queries += info.n_reads; pairs_done += (info.n_reads + 1) / 2;
          slot = (slot + 1) % SMASH_N_SLOTS;
          if (!eof && info.consumed[0] == 0 && len == cap) {          // a single line longer than the buffer: grow it
            char *bigger = (char *)smash_host_alloc(2 * cap);
            if (!bigger) throw std::runtime_error("out of pinned host memory");
            memcpy(bigger, text, len); smash_host_free(text); text = bigger; cap *= 2;
          }
          len -= info.consumed[0];
          memmove(text, text + info.consumed[0], len);
        }
        fclose(f);
        smash_host_free(text);
        more = false;
      }
      QueryParser *qpp = (device_reader || o.fastq_pair) ? nullptr : new QueryParser(o.inputs[fi], o);
      Batch buf[SMASH_N_SLOTS];
      while (more) {
        QueryParser &qp = *qpp;
        if (in_flight[slot]) drain(slot);
        Batch &b = buf[slot]; b.clear();
        while (b.n() < BATCH && (more = qp.next(b))) {}
        if (b.n()) {
          smash_batch v = b.view(pairs_done);
          check(smash_submit(ctx, slot, &v, want_sam));      // the other slot's batch is still on the GPU
          in_flight[slot] = true; n_queries += b.n(); pairs_done += (b.n() + 1) / 2;
          slot = (slot + 1) % SMASH_N_SLOTS;
        }
      }
      delete qpp;
      for (int s = 0; s < SMASH_N_SLOTS; ++s) { const int k = (slot + s) % SMASH_N_SLOTS; if (in_flight[k]) drain(k); }
      if (o.verbose && n_gpus == 1) std::cerr << "# query reader for " << o.inputs[fi] << " processed " << n_queries << " sequences" << std::endl;
    }
    total_queries += n_queries;
    };                                                               // run_gpu
    // the stages after mummer (-bins): collective over the GPUs, every rank ends up with the global counts
    std::vector<int64_t> counts(bin_starts.size());
    smash_tail_stats tstats{};
    auto finish_gpu = [&](int g) {
      if (o.bins.empty()) return;
      smash_tail_stats st_g{};
      std::vector<int64_t> mine(bin_starts.size());
      check(smash_bins_finish(ctxs[g], (uint64_t)g << 40, mine.data(), nullptr, &st_g));
      if (g == 0) { counts.swap(mine); tstats = st_g; }
    };
    if (n_gpus == 1) { run_gpu(0); finish_gpu(0); }
    else {
      std::vector<std::thread> th;
      std::vector<std::string> errs(n_gpus);
      for (int g = 0; g < n_gpus; ++g)
        th.emplace_back([&, g] {
          try { run_gpu(g); } catch (const std::exception &e) { errs[g] = e.what(); }
          try { if (errs[g].empty()) finish_gpu(g); } catch (const std::exception &e) { errs[g] = e.what(); }
        });
      for (auto &t : th) t.join();
      for (auto &e : errs) if (!e.empty()) throw std::runtime_error(e);
    }
    const uint64_t n_queries = total_queries.load();
    if (!o.bins.empty()) write_varbin(o, bin_rows, counts, tstats);
    if (!o.gc.empty()) write_gcnorm(o, bin_rows, counts);
    if (!n_queries) std::cerr << "# no reads processed" << std::endl;
    if (o.verbose)
      std::cerr << "# ran " << n_queries << " queries in "
                << std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - tq).count() << " seconds" << std::endl;
    for (size_t g = 1; g < ctxs.size(); ++g) smash_ctx_destroy(ctxs[g]);
    smash_ctx_destroy(ctx); smash_index_close(ix);
    return 0;
  } catch (const std::exception &e) {
    const std::string what = e.what();
    if (what.rfind("unable to open", 0) == 0) { std::cerr << what << std::endl; return 1; }   // reader thread path, query.cpp:689-695
    std::cerr << "Error" << std::endl << what << std::endl;                                   // mummer.cpp:61-65
    return 1;
  }
}
