// records.cpp -- see records.hpp. Every function cites the Java it mirrors (paths relative to src/main/java/edu/unc/bioinf/ubu/).
#include "records.hpp"

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "aligner.hpp"

namespace ubu {
namespace {

std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::ostringstream ss; ss << f.rdbuf();
  return ss.str();
}
void spit(const std::string& path, const std::string& text) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw std::runtime_error("cannot write " + path);
  f << text;
}
std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out; size_t a = 0;
  for (;;) { const size_t b = s.find(sep, a); if (b == std::string::npos) { out.push_back(s.substr(a)); break; } out.push_back(s.substr(a, b - a)); a = b + 1; }
  return out;
}
// String.split(sep): trailing empty strings are removed ("a;b;" -> {"a", "b"})
std::vector<std::string> javaSplit(const std::string& s, char sep) {
  std::vector<std::string> out = split(s, sep);
  while (!out.empty() && out.back().empty()) out.pop_back();
  return out;
}
bool endsWith(const std::string& s, const std::string& suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
// sam/ReverseComplementor.java:38-62: A<->T, C<->G, every other character unchanged
std::string reverseComplement(const std::string& s) {
  std::string out(s.rbegin(), s.rend());
  for (char& c : out) { switch (c) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; default: break; } }
  return out;
}
std::string reverse(const std::string& s) { return std::string(s.rbegin(), s.rend()); }

// ReadBlock.updateBasePositions (ReadBlock.java:134-156)
void updateBasePositions(int& readBase, int& refBase, char op, int n) {
  switch (op) {
    case 'S': readBase += n; break;
    case 'N': refBase += n; break;
    case 'D': refBase += n; break;
    case 'I': readBase += n; break;
    case 'M': readBase += n; refBase += n; break;
    default: throw std::logic_error(std::string("Case statement didn't deal with cigar op: ") + op);
  }
}
// ReadBlock.getSubBlock (ReadBlock.java:74-88)
ReadBlock getSubBlock(const ReadBlock& b, int accumulatedLength, int positionInRead, int maxLength) {
  const int positionInBlock = positionInRead + accumulatedLength - b.readStart + 1;
  if (b.type == 'N' || b.type == 'D') return ReadBlock{accumulatedLength + 1, b.referenceStart + positionInBlock, b.length - positionInBlock, b.type};
  if (b.type == 'S') return ReadBlock{accumulatedLength + 1, b.referenceStart, std::min(maxLength, b.length - positionInBlock), b.type};
  return ReadBlock{accumulatedLength + 1, b.referenceStart + positionInBlock, std::min(maxLength, b.length - positionInBlock), b.type};
}
int getTotalLength(const std::vector<ReadBlock>& blocks) {        // ReadBlock.java:101-111
  int n = 0; for (const ReadBlock& b : blocks) if (b.type != 'D') n += b.length; return n;
}
void fillToLength(std::vector<ReadBlock>& blocks, int readLength) {   // ReadBlock.java:158-175
  const int cigarLength = getTotalLength(blocks);
  if (cigarLength < readLength) {
    ReadBlock& last = blocks.back();
    if (last.type == 'M' || last.type == 'S') last.length += readLength - cigarLength;
    else {
      int rs = last.readStart, fs = last.referenceStart;
      updateBasePositions(rs, fs, last.type, last.length);
      blocks.push_back(ReadBlock{rs, fs, readLength - cigarLength, 'M'});
    }
  }
}
int getTotalLength(const std::vector<CigarElement>& e) {          // CombineChimera3.java:203-215
  int n = 0; for (const CigarElement& c : e) if (c.op != 'D') n += c.length; return n;
}

// java.util.HashMap<String, V>.values() order for the JVMs the reference targets (pom.xml: 1.6): table of 16 doubling at load factor
// 0.75, supplemental hash, buckets walked upwards, a bucket's chain newest first. One entry (the common case) is trivially itself.
uint32_t javaStringHash(const std::string& s) { uint32_t h = 0; for (unsigned char c : s) h = 31u * h + c; return h; }
std::vector<size_t> hashMapOrder(const std::vector<std::string>& keys) {
  std::vector<size_t> order;
  if (keys.size() <= 1) { for (size_t k = 0; k < keys.size(); ++k) order.push_back(k); return order; }
  size_t cap = 16; while ((double)keys.size() > cap * 0.75) cap *= 2;
  std::map<uint32_t, std::vector<size_t>> buckets;
  for (size_t k = 0; k < keys.size(); ++k) {
    uint32_t h = javaStringHash(keys[k]);
    h ^= (h >> 20) ^ (h >> 12); h = h ^ (h >> 7) ^ (h >> 4);
    std::vector<size_t>& b = buckets[h & (uint32_t)(cap - 1)];
    b.insert(b.begin(), k);
  }
  for (auto& kv : buckets) for (size_t k : kv.second) order.push_back(k);
  return order;
}

}  // namespace

// ---- SamRecord ----------------------------------------------------------------------------------------------------------------
std::vector<CigarElement> parseCigar(const std::string& text) {
  std::vector<CigarElement> out;
  if (text == "*") return out;
  int n = 0;
  for (char ch : text) { if (ch >= '0' && ch <= '9') n = n * 10 + (ch - '0'); else { out.push_back(CigarElement{n, ch}); n = 0; } }
  return out;
}

SamRecord SamRecord::parse(const std::string& lineIn) {
  std::string line = lineIn;
  while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
  const std::vector<std::string> f = split(line, '\t');
  if (f.size() < 11) throw std::runtime_error("malformed SAM line: " + line.substr(0, 80));
  SamRecord r;
  r.qname = f[0]; r.flag = std::stoi(f[1]); r.rname = f[2]; r.pos = std::stoi(f[3]); r.mapq = std::stoi(f[4]);
  r.cigar = parseCigar(f[5]);
  r.rnext = f[6] == "=" ? r.rname : f[6]; r.pnext = std::stoi(f[7]); r.tlen = std::stoi(f[8]);
  r.seq = f[9] == "*" ? "" : f[9]; r.qual = f[10] == "*" ? "" : f[10];
  for (size_t k = 11; k < f.size(); ++k) {
    const size_t a = f[k].find(':'), b = a == std::string::npos ? a : f[k].find(':', a + 1);
    if (b == std::string::npos) continue;
    r.tags.push_back(Tag{f[k].substr(0, a), f[k].substr(a + 1, b - a - 1), f[k].substr(b + 1)});
  }
  return r;
}

std::string SamRecord::getCigarString() const {
  if (cigar.empty()) return "*";
  std::string s;
  for (const CigarElement& e : cigar) { s += std::to_string(e.length); s += e.op; }
  return s;
}
int SamRecord::getCigarReadLength() const {
  int n = 0; for (const CigarElement& e : cigar) if (e.op == 'M' || e.op == 'I' || e.op == 'S' || e.op == '=' || e.op == 'X') n += e.length; return n;
}
int SamRecord::getAlignmentEnd() const {
  if (getReadUnmappedFlag()) return 0;
  int n = 0; for (const CigarElement& e : cigar) if (e.op == 'M' || e.op == 'D' || e.op == 'N' || e.op == '=' || e.op == 'X') n += e.length;
  return pos + n - 1;
}
bool SamRecord::getIntegerAttribute(const std::string& tag, int& out) const {
  for (const Tag& t : tags) if (t.tag == tag) { out = std::stoi(t.value); return true; }
  return false;
}
const std::string* SamRecord::getAttribute(const std::string& tag) const {
  for (const Tag& t : tags) if (t.tag == tag) return &t.value;
  return nullptr;
}
void SamRecord::setAttribute(const std::string& tag, int value) {
  for (Tag& t : tags) if (t.tag == tag) { t.type = "i"; t.value = std::to_string(value); return; }
  tags.push_back(Tag{tag, "i", std::to_string(value)});
}
void SamRecord::setAttribute(const std::string& tag, const std::string& value) {
  for (Tag& t : tags) if (t.tag == tag) { t.type = "Z"; t.value = value; return; }
  tags.push_back(Tag{tag, "Z", value});
}
void SamRecord::removeAttribute(const std::string& tag) {
  for (size_t k = 0; k < tags.size(); ++k) if (tags[k].tag == tag) { tags.erase(tags.begin() + k); return; }
}
std::string SamRecord::getSAMString() const {
  std::string s = qname; s += '\t'; s += std::to_string(flag); s += '\t'; s += rname; s += '\t'; s += std::to_string(pos); s += '\t';
  s += std::to_string(mapq); s += '\t'; s += getCigarString(); s += '\t';
  s += (rnext == rname && rnext != "*") ? "=" : rnext; s += '\t'; s += std::to_string(pnext); s += '\t'; s += std::to_string(tlen); s += '\t';
  s += getReadString(); s += '\t'; s += qual.empty() ? "*" : qual;
  for (const Tag& t : tags) { s += '\t'; s += t.tag; s += ':'; s += t.type; s += ':'; s += t.value; }
  s += '\n';
  return s;
}

void parseSamText(const std::string& text, std::vector<std::string>& header, std::vector<SamRecord>& records) {
  size_t a = 0;
  while (a < text.size()) {
    size_t b = text.find('\n', a);
    if (b == std::string::npos) b = text.size();
    if (b > a) {
      const std::string line = text.substr(a, b - a);
      if (line[0] == '@') header.push_back(line); else records.push_back(SamRecord::parse(line));
    }
    a = b + 1;
  }
}

// ---- ReadBlock / updateReadAlignment ------------------------------------------------------------------------------------------
std::vector<ReadBlock> ReadBlock::getReadBlocks(int alignmentStart, const std::vector<CigarElement>& cigar) {   // ReadBlock.java:114-132
  std::vector<ReadBlock> out;
  int readBase = 1, refBase = alignmentStart;
  for (const CigarElement& e : cigar) {
    out.push_back(ReadBlock{readBase, refBase, e.length, e.op});
    updateBasePositions(readBase, refBase, e.op, e.length);
  }
  return out;
}

bool updateReadAlignment(int contigPos, const std::vector<CigarElement>& contigCigar, int position, int readLength, int& newStart,
                         std::vector<CigarElement>& newCigar) {        // ReAligner.java:1293-1368
  std::vector<ReadBlock> blocks;
  int contigPosition = position, accumulatedLength = 0;
  for (const ReadBlock& contigBlock : ReadBlock::getReadBlocks(contigPos, contigCigar)) {
    if (contigBlock.readStart + contigBlock.getReferenceLength() >= position + 1) {              // :1306-1307
      ReadBlock block = getSubBlock(contigBlock, accumulatedLength, contigPosition, readLength - accumulatedLength);
      if (block.type == 'I' && block.length != 0) {                                              // :1314-1316
        contigPosition = contigPosition - (contigBlock.length - block.length);
        block.referenceStart = block.referenceStart - (contigBlock.length - block.length);
      }
      if (block.length != 0) {                                                                   // :1325
        blocks.push_back(block);
        if (block.type != 'D') accumulatedLength += block.length;
        if (accumulatedLength > readLength)
          throw std::logic_error("Accumulated Length: " + std::to_string(accumulatedLength) + " is greater than read length: " + std::to_string(readLength));
        if (accumulatedLength == readLength) break;
      }
    }
  }
  if (blocks.empty()) return false;                                                              // :1362-1364
  fillToLength(blocks, readLength);                                                              // :1356
  newStart = blocks[0].referenceStart;
  newCigar.clear();
  for (const ReadBlock& b : blocks) newCigar.push_back(CigarElement{b.length, b.type});          // toCigarString: blocks as they are
  return true;
}

// ---- CompareToReference.numDifferences (CompareToReference.java:83-123) ---------------------------------------------------------------
int numDifferences(const std::string& readBases, const std::string& refBases, int alignmentStart, const std::vector<CigarElement>& cigar) {
  int diffs = 0; size_t readIdx = 0, refIdx = (size_t)(alignmentStart - 1);
  for (const CigarElement& e : cigar) {
    if (e.op == 'M') {
      for (int i = 0; i < e.length; ++i) {
        if (readIdx >= readBases.size() || refIdx >= refBases.size()) throw std::out_of_range("String index out of range");
        const char a = (char)toupper((unsigned char)readBases[readIdx]), b = (char)toupper((unsigned char)refBases[refIdx]);
        if (a != b && a != 'N' && b != 'N') ++diffs;                                             // :92
        ++readIdx; ++refIdx;
      }
    } else if (e.op == 'I') readIdx += (size_t)e.length;
    else if (e.op == 'D') refIdx += (size_t)e.length;
    else if (e.op == 'S') readIdx += (size_t)e.length;                                           // N elements advance nothing (:99-108)
  }
  return diffs;
}
size_t countNonAcgtn(const std::string& bases) {
  size_t n = 0;
  for (unsigned char c : bases) { const char u = (char)toupper(c); if (u != 'A' && u != 'C' && u != 'G' && u != 'T' && u != 'N') ++n; }
  return n;
}

// ---- CombineChimera3 ------------------------------------------------------------------------------------------------------------
int CombineChimera3::getMappedReadStart(const SamRecord& read) {                                  // CombineChimera3.java:269-281
  for (const ReadBlock& b : ReadBlock::getReadBlocks(read.pos, read.cigar)) if (b.type != 'S') return b.readStart;
  return -1;
}
int CombineChimera3::getMappedReadEnd(const SamRecord& read) {                                    // CombineChimera3.java:283-296
  const std::vector<ReadBlock> blocks = ReadBlock::getReadBlocks(read.pos, read.cigar);
  for (size_t k = blocks.size(); k-- > 0;) if (blocks[k].type != 'S') return blocks[k].readStart + blocks[k].length - 1;
  return -1;
}
void CombineChimera3::sortReadsByPosition(std::vector<SamRecord>& reads) {                         // :221-237 (read2 - read1: descending; Collections.sort is stable)
  std::stable_sort(reads.begin(), reads.end(), [](const SamRecord& a, const SamRecord& b) {
    int cmp = getMappedReadStart(b) - getMappedReadStart(a);
    if (cmp == 0) cmp = getMappedReadEnd(b) - getMappedReadEnd(a);
    return cmp < 0;
  });
}
void CombineChimera3::pruneLikelyInserts(std::vector<SamRecord>& sortedReads) {                    // :251-267
  if (sortedReads.size() == 3) {
    const SamRecord &first = sortedReads[0], &middle = sortedReads[1], &last = sortedReads[2];
    const int FUDGE_FACTOR = 5;
    if (getMappedReadStart(middle) <= getMappedReadEnd(first) + FUDGE_FACTOR && getMappedReadEnd(middle) >= getMappedReadStart(last) - FUDGE_FACTOR)
      sortedReads.erase(sortedReads.begin() + 1);
  }
}
bool CombineChimera3::combineChimericReads(const SamRecord& read1, const SamRecord& read2, SamRecord& combined) {   // :75-184
  const SamRecord& left = read1.pos < read2.pos ? read1 : read2;
  const SamRecord& right = read1.pos < read2.pos ? read2 : read1;
  const std::vector<CigarElement>&leftCigar = left.cigar, &rightCigar = right.cigar;
  if (rightCigar[0].op == 'S' && leftCigar[leftCigar.size() - 1].op == 'S') {
    std::vector<CigarElement> leftElements(leftCigar.begin(), leftCigar.end() - 1);              // drop trailing S on the left side
    std::vector<CigarElement> rightElements(rightCigar.begin() + 1, rightCigar.end());           // drop leading S on the right side
    const std::vector<ReadBlock> leftBlocks = ReadBlock::getReadBlocks(left.pos, left.cigar);
    const ReadBlock& lastLeftBlock = leftBlocks[leftBlocks.size() - 2];
    const std::vector<ReadBlock> rightBlocks = ReadBlock::getReadBlocks(right.pos, right.cigar);
    const ReadBlock& firstRightBlock = rightBlocks[1];
    const int leftStop = lastLeftBlock.getReferenceStop(), rightStart = firstRightBlock.referenceStart;
    const int leftReadStop = lastLeftBlock.readStart + lastLeftBlock.length - 1, rightReadStart = firstRightBlock.readStart;
    int trimLength = 0;
    if (leftStop >= rightStart || leftReadStop >= rightReadStart) {
      trimLength = std::max(leftStop - rightStart + 1, leftReadStop - rightReadStart + 1);
      CigarElement& toTrim = leftElements[leftElements.size() - 1];                              // trimRightmostElement (:194-201)
      toTrim.length -= trimLength;
      if (toTrim.length < 1) return false;                                                       // not an indel, alternate mappings
    }
    const int leftAlignmentStop = lastLeftBlock.getReferenceStop() - trimLength;
    const int alignmentGap = firstRightBlock.referenceStart - leftAlignmentStop - 1;
    CigarElement indel;
    if (alignmentGap > 0) indel = CigarElement{alignmentGap, 'D'};
    else indel = CigarElement{read1.getReadLength() - (getTotalLength(leftElements) + getTotalLength(rightElements)), 'I'};
    combined = read1;                                                                            // cloneRead(read1)
    combined.pos = leftBlocks[0].referenceStart;
    combined.cigar = leftElements; combined.cigar.push_back(indel);
    combined.cigar.insert(combined.cigar.end(), rightElements.begin(), rightElements.end());
    combined.mapq = std::min(read1.mapq, read2.mapq);
    return true;
  }
  return false;
}
std::string CombineChimera3::combineText(const std::string& samText) {                             // :21-73
  std::vector<std::string> header; std::vector<SamRecord> recs;
  parseSamText(samText, header, recs);
  std::string out;
  for (std::string h : header) {
    if (h.compare(0, 3, "@HD") == 0) {                                                           // header.setSortOrder(SortOrder.unsorted)
      std::string nh;
      for (const std::string& f : split(h, '\t')) if (f.compare(0, 3, "SO:") != 0) { if (!nh.empty()) nh += '\t'; nh += f; }
      h = nh + "\tSO:unsorted";
    }
    out += h; out += '\n';
  }
  for (size_t k = 0; k < recs.size();) {                                                          // SamMultiMappingReader: consecutive records of one name
    size_t j = k;
    while (j < recs.size() && recs[j].qname == recs[k].qname) ++j;
    std::vector<SamRecord> readList(recs.begin() + k, recs.begin() + j);
    k = j;
    bool isCombined = false;
    sortReadsByPosition(readList);
    pruneLikelyInserts(readList);
    if (readList.size() == 2) {
      const SamRecord &read1 = readList[0], &read2 = readList[1];
      if (read1.rname == read2.rname && read1.getReadNegativeStrandFlag() == read2.getReadNegativeStrandFlag() && read1.cigar.size() >= 2 &&
          read2.cigar.size() >= 2) {
        SamRecord combined;
        if (combineChimericReads(read1, read2, combined)) { isCombined = true; out += combined.getSAMString(); }
      }
    }
    if (!isCombined) for (const SamRecord& r : readList) out += r.getSAMString();
  }
  return out;
}
void CombineChimera3::combine(const std::string& input, const std::string& output) { spit(output, combineText(slurp(input))); }

// ---- Sam2Fastq ------------------------------------------------------------------------------------------------------------------
namespace {
void appendFastq(std::string& out, const SamRecord& read) {                                       // samReadToFastqRecord (Sam2Fastq.java:111-162, non-fusion branch)
  std::string bases = read.getReadString(), qualities = read.qual.empty() ? "*" : read.qual;
  if (read.getReadNegativeStrandFlag()) { bases = reverseComplement(bases); qualities = reverse(qualities); }
  out += '@'; out += read.qname; out += '\n'; out += bases; out += "\n+\n"; out += qualities; out += '\n';
}
}  // namespace
void Sam2Fastq::setEndSuffixes(const std::string& e1, const std::string& e2) { byReadId_ = true; end1_ = e1; end2_ = e2; }
std::string Sam2Fastq::convertText(const std::string& samText) const {                           // Sam2Fastq.java:85-109
  std::vector<std::string> header; std::vector<SamRecord> recs;
  parseSamText(samText, header, recs);
  std::string out, last1Read;
  for (const SamRecord& read : recs) if (read.qname != last1Read) { appendFastq(out, read); last1Read = read.qname; }
  return out;
}
void Sam2Fastq::convertPairedText(const std::string& samText, std::string& out1, std::string& out2) const {   // Sam2Fastq.java:33-79
  std::vector<std::string> header; std::vector<SamRecord> recs;
  parseSamText(samText, header, recs);
  std::string last1Read, last2Read;
  int output1Count = 0, output2Count = 0;
  out1.clear(); out2.clear();
  for (const SamRecord& read : recs) {
    const bool first = byReadId_ ? endsWith(read.qname, end1_) : read.getFirstOfPairFlag();      // :164-188
    const bool second = byReadId_ ? endsWith(read.qname, end2_) : read.getSecondOfPairFlag();
    if (first) { if (read.qname != last1Read) { appendFastq(out1, read); last1Read = read.qname; ++output1Count; } }
    else if (second) { if (read.qname != last2Read) { appendFastq(out2, read); last2Read = read.qname; ++output2Count; } }
  }
  if (output1Count != output2Count) throw std::logic_error("Non-symmetrical read counts found.  Your reads may not be paired properly.");
}
void Sam2Fastq::convert(const std::string& inputSam, const std::string& outputFastq) { spit(outputFastq, convertText(slurp(inputSam))); }
void Sam2Fastq::convert(const std::string& inputSam, const std::string& o1, const std::string& o2) {
  std::string a, b; convertPairedText(slurp(inputSam), a, b); spit(o1, a); spit(o2, b);
}

// ---- ReAligner ------------------------------------------------------------------------------------------------------------------
void ReAligner::removeSoftClips(SamRecord& read) {                                                // ReAligner.java:926-961
  const std::vector<CigarElement>& cigar = read.cigar;
  const CigarElement firstElement = cigar[0], lastElement = cigar[cigar.size() - 1];
  if (firstElement.op == 'S' || lastElement.op == 'S') {
    std::vector<CigarElement> newCigar;
    std::string bases = read.getReadString();
    auto substring = [](const std::string& s, long a, long b) -> std::string {                   // String.substring(a, b)
      if (a < 0 || b > (long)s.size() || a > b) throw std::out_of_range("String index out of range");
      return s.substr((size_t)a, (size_t)(b - a));
    };
    if (firstElement.op == 'S') bases = substring(bases, firstElement.length, (long)bases.size() - 1);     // :942: the last base goes too
    else newCigar.push_back(firstElement);
    for (size_t i = 1; i + 1 < cigar.size(); ++i) newCigar.push_back(cigar[i]);
    if (lastElement.op == 'S') bases = substring(bases, 0, (long)bases.size() - lastElement.length - 1);   // :953: one base more than the clip
    read.cigar = newCigar;
    read.seq = bases == "*" ? "" : bases;
  }
}
std::string ReAligner::cleanAndOutputContigsText(const std::string& samText, int minContigLength, bool shouldRemoveSoftClips, bool& hasCleanContigs) {   // :876-921
  std::vector<std::string> header; std::vector<SamRecord> recs;
  parseSamText(samText, header, recs);
  std::string out;
  hasCleanContigs = false;
  for (SamRecord& contigRead : recs) {
    if (contigRead.mapq > 1) {                                                                   // :887
      if (shouldRemoveSoftClips) removeSoftClips(contigRead);
      const std::string bases = contigRead.getReadString();
      if ((int)bases.size() >= minContigLength) {                                                // :895
        contigRead.seq.clear();                                                                  // setReadString("")
        std::string contigReadStr = contigRead.getSAMString();
        std::replace(contigReadStr.begin(), contigReadStr.end(), '\t', '~');                     // the trailing '\n' of getSAMString stays (:904-905)
        out += ">" + contigRead.qname + "~" + contigReadStr;                                     // :907-909
        out += bases; out += "\n";
        hasCleanContigs = true;
      }
    }
  }
  return out;
}
bool ReAligner::cleanAndOutputContigs(const std::string& contigsSam, const std::string& cleanContigsFasta, bool shouldRemoveSoftClips) {
  bool has = false;
  spit(cleanContigsFasta, cleanAndOutputContigsText(slurp(contigsSam), minContigLength_, shouldRemoveSoftClips, has));
  return has;
}
std::string ReAligner::alignAndCleanContigs(const std::string& contigFasta, const std::string& tempDir, bool shouldRemoveSoftClips) {   // :351-367
  Aligner aligner(reference_, numThreads_, device_);
  const std::string contigsSam = tempDir + "/all_contigs.sam";
  aligner.align(contigFasta, contigsSam);
  CombineChimera3 cc;
  const std::string contigsWithChim = tempDir + "/all_contigs_chim.sam";
  cc.combine(contigsSam, contigsWithChim);
  const std::string cleanContigsFasta = tempDir + "/clean_contigs.fasta";
  return cleanAndOutputContigs(contigsWithChim, cleanContigsFasta, shouldRemoveSoftClips) ? cleanContigsFasta : std::string();
}
void ReAligner::alignToContigs(const std::string& tempDir, const std::string& inputSam, const std::string& alignedToContigSam,
                               const std::string& contigFasta) {                                  // :963-975
  const std::string fastq = tempDir + "/original_reads.fastq";
  Sam2Fastq().convert(inputSam, fastq);
  Aligner contigAligner(contigFasta, numThreads_, device_);
  contigAligner.index();
  contigAligner.shortAlign(fastq, alignedToContigSam);
}

namespace {
int getIntAttribute(const SamRecord& read, const char* tag) { int v = 0; return read.getIntegerAttribute(tag, v) ? v : 0; }        // :1028-1036
int getEditDistance(const SamRecord& read) { int v = 0; return read.getIntegerAttribute("NM", v) ? v : read.getReadLength(); }    // :1038-1046
struct HitInfo { SamRecord record; int position; char strand; int mismatches; };                                                // :990-1026
SamRecord contigRecordFromName(const std::string& name) {          // substring(indexOf('~') + 1).replace('~', '\t') -> SamStringReader.getRead (:1102-1104)
  std::string s = name.substr(name.find('~') == std::string::npos ? 0 : name.find('~') + 1);
  std::replace(s.begin(), s.end(), '~', '\t');
  return SamRecord::parse(s);
}
}  // namespace

std::string ReAligner::adjustReadsText(const std::string& origSam, const std::string& toContigSam, const std::string& fullMatchCigarIn, AdjustStats& stats) {   // :1048-1291
  std::vector<std::string> h1, h2; std::vector<SamRecord> origs, contigs;
  parseSamText(origSam, h1, origs); parseSamText(toContigSam, h2, contigs);
  std::string out;
  size_t ci = 0, oi = 0; bool haveCached = false; SamRecord cachedContig;
  while ((ci < contigs.size() || haveCached) && oi < origs.size()) {                               // :1070
    const SamRecord& orig = origs[oi++];
    SamRecord read;
    if (haveCached) { read = cachedContig; haveCached = false; } else read = contigs[ci++];
    bool isReadWritten = false;
    if (orig.qname == read.qname) {
      const std::string fullMatchCigar = fullMatchCigarIn.empty() ? std::to_string(read.getReadLength()) + "M" : fullMatchCigarIn;
      if (read.getCigarString() == fullMatchCigar && !read.getReadUnmappedFlag() && getEditDistance(read) < getEditDistance(orig)) {   // :1091-1093
        std::vector<HitInfo> bestHits;
        SamRecord contigRead = contigRecordFromName(read.rname);
        const int bestMismatches = getIntAttribute(read, "XM");
        if (read.getAlignmentEnd() <= contigRead.getCigarReadLength())                             // :1112
          bestHits.push_back(HitInfo{contigRead, read.pos, read.getReadNegativeStrandFlag() ? '-' : '+', bestMismatches});
        const int numBestHits = getIntAttribute(read, "X0"), subOptimalHits = getIntAttribute(read, "X1");
        const int totalHits = numBestHits + subOptimalHits;
        if (totalHits > 1 && totalHits < 1000) {                                                   // :1126
          const std::string* alternateHitsStr = read.getAttribute("XA");
          if (!alternateHitsStr) stats.missingXATag += 1;
          else {
            const std::vector<std::string> alternates = javaSplit(*alternateHitsStr, ';');
            for (size_t i = 0; i + 1 < alternates.size(); ++i) {                                   // :1137 (i < length-1: the last entry is skipped)
              const std::vector<std::string> altInfo = split(alternates[i], ',');
              const char strand = altInfo.at(1).at(0);
              const int position = std::stoi(altInfo.at(1).substr(1));
              const std::string& cigar = altInfo.at(2);
              const int mismatches = std::stoi(altInfo.at(3));
              if (cigar == fullMatchCigar && mismatches < bestMismatches) stats.mismatchIssue += 1;
              if (cigar == fullMatchCigar && mismatches == bestMismatches) {
                SamRecord alt = contigRecordFromName(altInfo.at(0));
                if (position + read.getReadLength() <= alt.getCigarReadLength()) bestHits.push_back(HitInfo{alt, position, strand, mismatches});   // :1156
              }
            }
          }
        }
        std::vector<std::string> keys; std::vector<SamRecord> values;                             // Map<String, SAMRecord> outputReadAlignmentInfo
        for (const HitInfo& hitInfo : bestHits) {
          const SamRecord& cr = hitInfo.record;
          const int position = hitInfo.position - 1;
          if (cr.mapq > orig.mapq) {                                                               // :1176
            int newStart = 0; std::vector<CigarElement> newCigar;
            if (updateReadAlignment(cr.pos, cr.cigar, position, orig.getReadLength(), newStart, newCigar)) {
              SamRecord updatedRead = orig;                                                        // cloneRead(orig.getRead())
              updatedRead.rname = cr.rname; updatedRead.cigar = newCigar; updatedRead.pos = newStart;
              if (updatedRead.mapq == 0) updatedRead.mapq = 1;
              updatedRead.flag &= ~0x4;
              const bool neg = hitInfo.strand == '-';
              updatedRead.flag = neg ? (updatedRead.flag | 0x10) : (updatedRead.flag & ~0x10);
              if (neg == read.getReadNegativeStrandFlag()) { updatedRead.seq = read.seq; updatedRead.qual = read.qual; }     // :1200-1205
              else { updatedRead.seq = reverseComplement(read.seq); updatedRead.qual = reverse(read.qual); }
              if (orig.getReadUnmappedFlag() || orig.rname != updatedRead.rname || orig.pos != updatedRead.pos ||
                  orig.getReadNegativeStrandFlag() != updatedRead.getReadNegativeStrandFlag() || orig.getCigarString() != updatedRead.getCigarString()) {
                std::string originalAlignment = "N/A";
                if (!orig.getReadUnmappedFlag())
                  originalAlignment = orig.rname + ":" + std::to_string(orig.pos) + ":" + (orig.getReadNegativeStrandFlag() ? "-" : "+") + ":" + orig.getCigarString();
                updatedRead.setAttribute("YO", originalAlignment);
              }
              updatedRead.setAttribute("YM", hitInfo.mismatches);
              updatedRead.setAttribute("YQ", cr.mapq);
              const std::string key = updatedRead.rname + "_" + std::to_string(updatedRead.pos) + "_" + (updatedRead.getReadNegativeStrandFlag() ? "-" : "+") +
                                      "_" + updatedRead.getCigarString();
              if (std::find(keys.begin(), keys.end(), key) == keys.end()) { keys.push_back(key); values.push_back(updatedRead); }
            }
          }
        }
        for (size_t k : hashMapOrder(keys)) {                                                      // outputReadAlignmentInfo.values() (:1241)
          SamRecord& readToOutput = values[k];
          const int origBestHits = getIntAttribute(readToOutput, "X0"), origSuboptimalHits = getIntAttribute(readToOutput, "X1");
          if (keys.size() > 1 || totalHits > 1000) readToOutput.mapq = 0;                          // :1249-1251
          if (readToOutput.getAttribute("YO")) {
            readToOutput.setAttribute("X0", (int)keys.size());
            readToOutput.setAttribute("X1", origBestHits + origSuboptimalHits);
            for (const char* t : {"XO", "XG", "MD", "XA", "XT"}) readToOutput.removeAttribute(t);
          }
          out += readToOutput.getSAMString();
          isReadWritten = true;
          stats.adjusted += 1;
        }
      }
    } else {
      cachedContig = read; haveCached = true;
    }
    if (!isReadWritten) out += orig.getSAMString();
  }
  return out;
}
void ReAligner::adjustReads(const std::string& originalReadsSam, const std::string& alignedToContigSam, const std::string& outputSam) {
  const std::string origText = slurp(originalReadsSam);
  std::vector<std::string> header; std::vector<SamRecord> none;
  std::string head;
  { size_t a = 0; while (a < origText.size() && origText[a] == '@') { const size_t b = origText.find('\n', a); head += origText.substr(a, b - a + 1); a = b + 1; } }
  lastAdjustStats = AdjustStats();
  spit(outputSam, head + adjustReadsText(origText, slurp(alignedToContigSam), fullMatchCigar, lastAdjustStats));
}

}  // namespace ubu
