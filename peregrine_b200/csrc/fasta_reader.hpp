// fasta_reader.hpp — host-side FASTA/FASTQ record scanner of shmr_mkseqdb (src/shmr_mkseqdb.c:99-121).
//
// The reference reads its inputs with klib's kseq (src/kseq.h:185-224, KSEQ_INIT(gzFile, gzread)); what reaches the .idx and
// .seqdb files is fixed by kseq's record grammar, restated here over a whole-file buffer:
//   * with no pending header character, everything up to the next '>' or '@' ANYWHERE in the text is skipped (:189-192);
//   * the name runs to the first isspace() character; unless that was '\n', the rest of the line is a comment (:195-196);
//   * sequence lines follow until a line STARTS with '>', '+' or '@' (:201-205); empty lines are skipped; after every line
//     one trailing '\r' is dropped if the accumulated sequence is longer than one character (kseq.h:138);
//   * '+' starts a FASTQ quality block: rest of that line skipped, then whole lines are appended until the quality is at
//     least as long as the sequence (:218-220); a length mismatch or a missing quality block ends the FILE (kseq_read
//     returns -2 and the caller's `while (kseq_read(seq) >= 0)` loop stops, src/shmr_mkseqdb.c:108);
//   * after a FASTQ record the scanner looks for the next header from scratch (last_char = 0, :221).
// Only the byte-level grammar lives here (plain C++, also compiled into tests/hostsim); encoding the bases is the GPU's job.
#pragma once
#include <algorithm>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>

namespace pgb {

struct FastaRecord {
  std::string name;
  std::string seq;  // ASCII, line breaks removed
};

class FastaScanner {
 public:
  FastaScanner(const char *buf, size_t n) : b_(buf), n_(n) {}
  // streaming use (GzRecordStream below): the scanner is a function of the byte stream and one remembered header character, so
  // a record can be re-scanned from its first byte once more of the file is in the buffer
  FastaScanner(const char *buf, size_t n, int last_char) : b_(buf), n_(n), last_char_(last_char) {}
  size_t pos() const { return pos_; }
  int last_char() const { return last_char_; }
  // next record into r; false at end of input or at a malformed FASTQ record (both end the file in the reference)
  bool next(FastaRecord &r) {
    if (last_char_ == 0) {
      while (pos_ < n_ && b_[pos_] != '>' && b_[pos_] != '@') pos_++;
      if (pos_ >= n_) return false;
      last_char_ = b_[pos_++];
    }
    if (pos_ >= n_) return false;  // ks_getuntil at end of file: normal exit
    r.name.clear();
    r.seq.clear();
    size_t i = pos_;
    while (i < n_ && !is_space(b_[i])) i++;
    r.name.assign(b_ + pos_, i - pos_);
    int c = i < n_ ? (unsigned char)b_[i] : 0;
    pos_ = i < n_ ? i + 1 : n_;
    if (c != '\n') skip_line();  // comment
    int c2 = -1;
    while (pos_ < n_) {
      c2 = (unsigned char)b_[pos_++];
      if (c2 == '>' || c2 == '+' || c2 == '@') break;
      if (c2 == '\n') { c2 = -1; continue; }
      r.seq.push_back((char)c2);
      c2 = -1;
      if (pos_ >= n_) break;  // ks_getuntil2 returns -1 before touching the string
      append_line(r.seq);
    }
    if (c2 == '>' || c2 == '@') last_char_ = c2;
    if (c2 != '+') return true;  // FASTA
    // FASTQ
    bool eol = false;
    while (pos_ < n_) if (b_[pos_++] == '\n') { eol = true; break; }
    if (!eol) return false;  // no quality string
    qual_.clear();
    while (pos_ < n_) {
      append_line(qual_);
      if (qual_.size() >= r.seq.size()) break;
    }
    last_char_ = 0;
    return qual_.size() == r.seq.size();
  }

 private:
  static bool is_space(char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\v' || ch == '\f' || ch == '\r'; }
  void skip_line() {
    const char *e = pos_ < n_ ? (const char *)memchr(b_ + pos_, '\n', n_ - pos_) : nullptr;
    pos_ = e ? (size_t)(e - b_) + 1 : n_;
  }
  // ks_getuntil2(KS_SEP_LINE, append = 1): rest of the current line, then kseq.h:138's '\r' rule on the whole string
  void append_line(std::string &s) {
    const char *e = (const char *)memchr(b_ + pos_, '\n', n_ - pos_);
    const size_t end = e ? (size_t)(e - b_) : n_;
    s.append(b_ + pos_, end - pos_);
    pos_ = e ? end + 1 : n_;
    if (s.size() > 1 && s.back() == '\r') s.pop_back();
  }
  const char *b_;
  size_t n_, pos_ = 0;
  int last_char_ = 0;
  std::string qual_;
};

// Records of one (optionally gzip-compressed) file without holding the file in memory: the text is read in blocks; a record whose
// scan touched the end of the buffered text while the file has more is scanned again after the next block arrived, so the
// buffer holds one block plus at most one record (the reference's kseq streams the same way, src/kseq.h:92-130).
class GzRecordStream {
 public:
  explicit GzRecordStream(const char *path, size_t block = (size_t)64 << 20) : block_(block) {
    f_ = gzopen(path, "r");
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~GzRecordStream() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != nullptr; }
  bool next(FastaRecord &r) {
    for (;;) {
      FastaScanner sc(buf_.data() + start_, buf_.size() - start_, last_char_);
      const bool got = sc.next(r);
      if (!eof_ && start_ + sc.pos() >= buf_.size()) {  // the scan ran into the end of what is buffered: more text may belong to it
        fill();
        continue;
      }
      if (!got) return false;  // end of file, or a malformed FASTQ record (both end the file in the reference)
      start_ += sc.pos();
      last_char_ = sc.last_char();
      return true;
    }
  }

 private:
  void fill() {
    if (start_) { buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)start_); start_ = 0; }
    const size_t old = buf_.size();
    buf_.resize(old + block_);
    size_t got_total = 0;
    while (got_total < block_) {
      const int got = gzread(f_, buf_.data() + old + got_total, (unsigned)std::min<size_t>(block_ - got_total, (size_t)1 << 30));
      if (got <= 0) { eof_ = true; break; }
      got_total += (size_t)got;
    }
    buf_.resize(old + got_total);
  }
  gzFile f_ = nullptr;
  size_t block_, start_ = 0;
  std::vector<char> buf_;
  int last_char_ = 0;
  bool eof_ = false;
};

// whole file into memory through zlib (plain and gzip-compressed files alike, as gzopen/gzread do for the reference)
inline bool slurp_gz(const char *path, std::vector<char> &out) {
  gzFile f = gzopen(path, "r");
  if (!f) return false;
  gzbuffer(f, 1 << 20);
  out.clear();
  size_t cap = (size_t)1 << 24;
  for (;;) {
    if (out.size() + ((size_t)1 << 22) > cap) cap *= 2;
    out.reserve(cap);
    const size_t old = out.size();
    out.resize(old + ((size_t)1 << 22));
    const int got = gzread(f, out.data() + old, 1 << 22);
    out.resize(old + (got > 0 ? (size_t)got : 0));
    if (got <= 0) break;
  }
  gzclose(f);
  return true;
}

}  // namespace pgb
