// aligner.cpp -- see aligner.hpp. Text I/O and bookkeeping only; the alignment runs in libubu_sw.so on the GPU.
#include "aligner.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <unordered_map>
#include <stdexcept>
#include <vector>

#include "../include/ubu_sw.h"

namespace ubu {
namespace {

struct Rec { std::string name, seq, qual; bool has_qual = false; };

std::string first_token(const std::string& s) {
  size_t a = 0; while (a < s.size() && isspace((unsigned char)s[a])) ++a;
  size_t b = a; while (b < s.size() && !isspace((unsigned char)s[b])) ++b;
  return s.substr(a, b - a);
}
void rstrip(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }

// FASTA, or FASTQ with 4-line records (as the reference's FastqInputFile reads them)
std::vector<Rec> read_fastx(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<Rec> out;
  std::string line;
  if (!std::getline(f, line)) return out;
  rstrip(line);
  if (!line.empty() && line[0] == '>') {
    Rec cur; cur.name = first_token(line.substr(1));
    while (std::getline(f, line)) {
      rstrip(line);
      if (!line.empty() && line[0] == '>') { out.push_back(cur); cur = Rec(); cur.name = first_token(line.substr(1)); }
      else { size_t a = 0, b = line.size(); while (a < b && isspace((unsigned char)line[a])) ++a; while (b > a && isspace((unsigned char)line[b - 1])) --b; cur.seq += line.substr(a, b - a); }
    }
    out.push_back(cur);
  } else if (!line.empty() && line[0] == '@') {
    std::string head = line;
    while (true) {
      Rec r; r.name = first_token(head.substr(1)); r.has_qual = true;
      std::string plus;
      if (!std::getline(f, r.seq)) r.seq.clear();
      rstrip(r.seq);
      std::getline(f, plus);
      if (!std::getline(f, r.qual)) r.qual.clear();
      rstrip(r.qual);
      out.push_back(r);
      if (!std::getline(f, head)) break;
      rstrip(head);
      if (head.empty()) break;
    }
  } else {
    throw std::runtime_error(path + ": neither FASTA nor FASTQ");
  }
  return out;
}

std::string revcomp(const std::string& s) {
  std::string o(s.size(), 'N');
  ubu_sw_revcomp(s.data(), s.size(), &o[0]);
  return o;
}

struct Packed {
  std::vector<uint8_t> packed, nmask; std::vector<uint32_t> off, noff; std::vector<uint16_t> len; bool any_n = false;
  void add(const std::string& s) {
    const size_t b0 = packed.size(), m0 = nmask.size();
    off.push_back((uint32_t)b0); noff.push_back((uint32_t)m0); len.push_back((uint16_t)s.size());
    packed.resize(b0 + (s.size() + 3) / 4); nmask.resize(m0 + (s.size() + 7) / 8);
    int hn = 0;
    if (!s.empty()) ubu_sw_pack(s.data(), s.size(), packed.data() + b0, nmask.data() + m0, &hn);
    any_n = any_n || hn;
  }
};

// exact k-mer index over the reference sequences (k <= 31, 2 bits per base, k-mers with a non-ACGT letter skipped)
struct SeedIndex {
  int k = 0; std::unordered_map<uint64_t, std::vector<uint32_t>> map;
  static int code(char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'T': case 't': return 2; case 'G': case 'g': return 3; default: return -1; } }
  template <class F> void each_kmer(const std::string& s, F f) const {
    const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    uint64_t v = 0; int run = 0;
    for (char ch : s) {
      const int c = code(ch);
      if (c < 0) { run = 0; v = 0; continue; }
      v = ((v << 2) | (uint64_t)c) & mask;
      if (++run >= k) f(v);
    }
  }
  void build(const std::vector<Rec>& refs, int kk) {
    k = kk;
    for (size_t w = 0; w < refs.size(); ++w)
      each_kmer(refs[w].seq, [&](uint64_t v) { std::vector<uint32_t>& l = map[v]; if (l.empty() || l.back() != (uint32_t)w) l.push_back((uint32_t)w); });
  }
  std::vector<uint32_t> candidates(const std::string& read) const {          // ascending window order
    std::vector<uint32_t> out;
    each_kmer(read, [&](uint64_t v) { auto it = map.find(v); if (it != map.end()) out.insert(out.end(), it->second.begin(), it->second.end()); });
    std::sort(out.begin(), out.end()); out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
  }
};

std::string cigar_text(const uint32_t* ops, uint32_t n) {
  char buf[4096];
  ubu_sw_cigar_string(ops, n, buf, sizeof buf);
  return buf;
}

}  // namespace

Aligner::Aligner(const std::string& reference, int numThreads, int device, int match, int mismatch, int gapOpen, int gapExt)
    : reference_(reference), numThreads_(numThreads), device_(device), match_(match), mismatch_(mismatch), gapOpen_(gapOpen), gapExt_(gapExt) {}

Aligner::~Aligner() { if (handle_) ubu_sw_destroy((ubu_sw_handle*)handle_); }

void Aligner::index() {
  std::ifstream f(reference_);
  if (!f) throw std::runtime_error("reference not found: " + reference_);
}
void Aligner::align(const std::string& input, const std::string& outputSam) { run(input, outputSam); }
void Aligner::shortAlign(const std::string& input, const std::string& outputSam) { run(input, outputSam); }

void Aligner::run(const std::string& input, const std::string& outputSam) {
  const std::vector<Rec> refs = read_fastx(reference_);
  const std::vector<Rec> reads = read_fastx(input);
  std::ofstream out(outputSam);
  if (!out) throw std::runtime_error("cannot write " + outputSam);
  int seed_k = seedK_;
  if (seed_k < 0) { const char* e = getenv("UBU_ALIGNER_SEED_K"); seed_k = e ? atoi(e) : 0; }
  seed_k = std::max(0, std::min(seed_k, 31));
  out << "@HD\tVN:1.4\tSO:unsorted\n";
  for (const Rec& r : refs) out << "@SQ\tSN:" << r.name << "\tLN:" << r.seq.size() << "\n";
  out << "@PG\tID:ubu_b200\tPN:ubu_b200\tDS:window Smith-Waterman on B200 (SPEC.md); X0 X1 XA MAPQ rules are BUILD-DEFINED (best = highest score, X1 = within one mismatch of it, MAPQ 37/23/0), not bwa's"
      << (seed_k > 0 ? "; candidates = sequences sharing an exact " + std::to_string(seed_k) + "-mer with the read" : std::string()) << "\n";
  auto unmapped = [&](const Rec& rd) {
    out << rd.name << "\t4\t*\t0\t0\t*\t*\t0\t0\t" << rd.seq << "\t" << (rd.has_qual && !rd.qual.empty() ? rd.qual : "*") << "\n";
  };
  if (reads.empty()) return;
  if (refs.empty()) { for (const Rec& rd : reads) unmapped(rd); return; }
  for (const Rec& r : refs) if (r.seq.size() > UBU_SW_MAX_WINDOW) throw std::runtime_error("reference sequence longer than 65535: " + r.name);
  for (const Rec& r : reads) if (r.seq.size() > UBU_SW_MAX_READ) throw std::runtime_error("read longer than 4096: " + r.name);

  if (!handle_) {
    ubu_sw_params p{match_, mismatch_, gapOpen_, gapExt_};
    ubu_sw_handle* h = nullptr;
    const int rc = ubu_sw_create(&p, device_, 0, &h);   // default pipeline depth: large batches stream through it inside align_batch
    if (rc) throw std::runtime_error(std::string("ubu_sw_create: ") + ubu_sw_strerror(rc));
    handle_ = h;
  }
  ubu_sw_handle* h = (ubu_sw_handle*)handle_;

  Packed W;
  for (const Rec& r : refs) W.add(r.seq);
  const size_t nw = refs.size();
  SeedIndex seeds;
  if (seed_k > 0) seeds.build(refs, seed_k);
  const size_t per_batch = seed_k > 0 ? 500000 : std::max<size_t>(1, 2000000 / std::max<size_t>(1, 2 * nw));
  for (size_t lo = 0; lo < reads.size(); lo += per_batch) {
    const size_t hi = std::min(reads.size(), lo + per_batch), nr = hi - lo;
    Packed R;                                    // entries: fwd_0, rc_0, fwd_1, rc_1, ...
    for (size_t k = lo; k < hi; ++k) { R.add(reads[k].seq); R.add(revcomp(reads[k].seq)); }
    // tasks: (entry e, window w) pairs sharing the packed bytes -- all of them, or the seeded candidates; entry_first[e] .. entry_first[e+1]
    // are entry e's tasks in window order
    std::vector<uint32_t> toff, tnoff, widx, entry_first(2 * nr + 1, 0); std::vector<uint16_t> tlen;
    for (size_t e = 0; e < 2 * nr; ++e) {
      entry_first[e] = (uint32_t)widx.size();
      std::vector<uint32_t> cand;
      if (seed_k > 0) cand = seeds.candidates(e % 2 ? revcomp(reads[lo + e / 2].seq) : reads[lo + e / 2].seq);
      else { cand.resize(nw); for (size_t w = 0; w < nw; ++w) cand[w] = (uint32_t)w; }
      for (uint32_t w : cand) { toff.push_back(R.off[e]); tnoff.push_back(R.noff[e]); tlen.push_back(R.len[e]); widx.push_back(w); }
    }
    entry_first[2 * nr] = (uint32_t)widx.size();
    const size_t nt = widx.size();
    if (nt == 0) { for (size_t k = lo; k < hi; ++k) unmapped(reads[k]); continue; }
    ubu_sw_seqs rs{}, ws{};
    rs.packed = R.packed.data(); rs.off = toff.data(); rs.len = tlen.data(); rs.count = (uint32_t)nt; rs.packed_bytes = (uint32_t)R.packed.size();
    if (R.any_n) { rs.nmask = R.nmask.data(); rs.nmask_off = tnoff.data(); rs.nmask_bytes = (uint32_t)R.nmask.size(); }
    ws.packed = W.packed.data(); ws.off = W.off.data(); ws.len = W.len.data(); ws.count = (uint32_t)nw; ws.packed_bytes = (uint32_t)W.packed.size();
    if (W.any_n) { ws.nmask = W.nmask.data(); ws.nmask_off = W.noff.data(); ws.nmask_bytes = (uint32_t)W.nmask.size(); }
    std::vector<ubu_sw_result> res(nt);
    std::vector<uint32_t> pool(6 * nt + 4096);
    size_t used = 0;
    int rc = ubu_sw_align_batch(h, &rs, &ws, widx.data(), res.data(), pool.data(), pool.size(), &used);
    if (rc == UBU_SW_ERR_CIGAR_POOL) { pool.resize(used + 16); rc = ubu_sw_align_batch(h, &rs, &ws, widx.data(), res.data(), pool.data(), pool.size(), &used); }
    if (rc) throw std::runtime_error(std::string("aligner failed: ") + ubu_sw_strerror(rc) + " (" + ubu_sw_last_error(h) + ")");

    for (size_t k = 0; k < nr; ++k) {
      const Rec& rd = reads[lo + k];
      // this read's tasks as (window, strand, task), reference order, plus strand first
      struct Cand { uint32_t w; int s; size_t t; };
      std::vector<Cand> cands;
      {
        size_t a = entry_first[2 * k], a1 = entry_first[2 * k + 1], b = a1, b1 = entry_first[2 * k + 2];
        while (a < a1 || b < b1) {
          if (b >= b1 || (a < a1 && widx[a] <= widx[b])) { cands.push_back({widx[a], 0, a}); ++a; } else { cands.push_back({widx[b], 1, b}); ++b; }
        }
      }
      std::unordered_map<uint64_t, size_t> task_of;
      for (const Cand& c : cands) task_of[((uint64_t)c.w << 1) | (uint64_t)c.s] = c.t;
      auto task = [&](size_t w, int s) { return task_of[((uint64_t)w << 1) | (uint64_t)s]; };
      int best = 0;
      for (const Cand& c : cands) best = std::max(best, res[c.t].score);
      if (best <= 0) { unmapped(rd); continue; }
      std::vector<std::pair<size_t, int>> hits;                   // reference order, plus strand first
      int x1 = 0;
      for (const Cand& c : cands) {
        const int sc = res[c.t].score;
        if (sc == best) hits.push_back({c.w, c.s});
        else if (sc >= best - (match_ - mismatch_) && sc < best) ++x1;
      }
      const size_t t0 = task(hits[0].first, hits[0].second);
      const ubu_sw_result& r0 = res[t0];
      const std::string cig = cigar_text(pool.data() + r0.cigar_off, r0.cigar_n);
      int gaps = 0;
      for (uint32_t c = 0; c < r0.cigar_n; ++c) { const uint32_t op = pool[r0.cigar_off + c] & 15u; if (op == 1 || op == 2) gaps += (int)(pool[r0.cigar_off + c] >> 4); }
      const int nm = r0.edit_distance, xm = nm - gaps;
      const int mapq = hits.size() > 1 ? 0 : (x1 ? 23 : 37);
      const bool minus = hits[0].second == 1;
      std::string qual = rd.has_qual && !rd.qual.empty() ? rd.qual : "*";
      if (minus && rd.has_qual && !rd.qual.empty()) std::reverse(qual.begin(), qual.end());
      out << rd.name << "\t" << (minus ? 16 : 0) << "\t" << refs[hits[0].first].name << "\t" << r0.ref_start << "\t" << mapq << "\t" << cig
          << "\t*\t0\t0\t" << (minus ? revcomp(rd.seq) : rd.seq) << "\t" << qual << "\tNM:i:" << nm << "\tXM:i:" << xm << "\tX0:i:" << hits.size()
          << "\tX1:i:" << x1 << "\tAS:i:" << best;
      if (hits.size() > 1) {
        out << "\tXA:Z:";
        for (size_t a = 1; a < hits.size() && a <= 1000; ++a) {     // `bwa samse -n 1000` (Aligner.java:53)
          const ubu_sw_result& rr = res[task(hits[a].first, hits[a].second)];
          out << refs[hits[a].first].name << "," << (hits[a].second ? '-' : '+') << rr.ref_start << "," << cigar_text(pool.data() + rr.cigar_off, rr.cigar_n)
              << "," << rr.edit_distance << ";";
        }
      }
      out << "\n";
    }
  }
}

}  // namespace ubu
