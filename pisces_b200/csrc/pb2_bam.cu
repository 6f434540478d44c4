// BAM -> struct-of-arrays read stager (SURVEY 8f rank 2, and rows a1-a4 of the path on the host side): BGZF inflate (zlib), record decode, the read
// filter, per-base directions from the XD tag and the collapsed-read summary from XV / XW / XR. Host code only; its output is exactly a pb2_read_batch.
//   BamReader.GetNextAlignment / decode          src/lib/Alignment.IO/BamReader.cs:137-224 (4-bit bases "=ACMGRSVTWYHKDBN", CIGAR count = low 16 bits of flag_nc)
//   AlignmentSource.ShouldSkipRead                src/exe/Pisces/Logic/Alignment/AlignmentsSource.cs:84-92
//   stitched / collapsed detection from @PG       src/lib/Pisces.IO/BamFileAlignmentExtractor.cs:111-153
//   Read.SequencedBaseDirectionMap                src/lib/Pisces.Domain/Models/Read.cs:390-421,664-682; CigarDirection (XD) Models/CigarDirection.cs:71-83
//   IsCollapsedRead / IsDuplex / ReadPairDirection Read.cs:66-71,311-349
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/pisces_b200.h"

struct pb2_bam_reader {
    FILE* f = nullptr;
    std::string error;
    std::vector<uint8_t> buf;   // inflated bytes not consumed yet
    size_t cur = 0;
    bool eof = false;
    std::string header_text;
    std::vector<std::string> ref_names;
    std::vector<const char*> ref_name_ptrs;
    std::vector<int32_t> ref_lengths;
    int32_t is_stitched = 0, is_collapsed = 0;
    // the batch handed out last
    std::vector<int32_t> pos0;
    std::vector<uint16_t> flag;
    std::vector<int64_t> cigar_off, seq_off;
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> bases, quals, base_dirs, collapsed;
    // the packed form of the batch handed out last (pb2_bam_next_batch_packed)
    std::vector<uint8_t> packed_seq, exc_base, exc_qual, cigar_ops;
    std::vector<int64_t> exc_index;
    bool batch_any_dirs = false, batch_any_coll = false;
    std::vector<int32_t> amplicon_id;                    // per read of the batch: index into amplicon_names, -1 without an XN tag
    std::vector<std::string> amplicon_names;             // the file's amplicon names in first-seen order (kept reads only)
    std::vector<const char*> amplicon_name_ptrs;
    // a record read ahead that belongs to the next chromosome
    std::vector<uint8_t> pending;
    bool have_pending = false;
};

namespace {
int bfail(pb2_bam_reader* r, const std::string& m) { if (r) r->error = m; return PB2_ERR_ARG; }

// inflate the next BGZF block into r->buf; false at end of file
bool next_block(pb2_bam_reader* r) {
    uint8_t hd[12];
    const size_t got = fread(hd, 1, 12, r->f);
    if (got == 0) { r->eof = true; return false; }
    if (got != 12 || hd[0] != 0x1f || hd[1] != 0x8b || hd[2] != 8 || !(hd[3] & 4)) { r->error = "not a BGZF block"; r->eof = true; return false; }
    const int xlen = hd[10] | (hd[11] << 8);
    std::vector<uint8_t> extra((size_t)xlen);
    if (fread(extra.data(), 1, (size_t)xlen, r->f) != (size_t)xlen) { r->error = "truncated BGZF header"; r->eof = true; return false; }
    int bsize = -1;
    for (int q = 0; q + 4 <= xlen;) {
        const int slen = extra[(size_t)q + 2] | (extra[(size_t)q + 3] << 8);
        if (extra[(size_t)q] == 66 && extra[(size_t)q + 1] == 67 && slen == 2) bsize = extra[(size_t)q + 4] | (extra[(size_t)q + 5] << 8);
        q += 4 + slen;
    }
    if (bsize < 0) { r->error = "BGZF block without a BC field"; r->eof = true; return false; }
    // BSIZE + 1 = the whole block: 12 header bytes, XLEN extra bytes, the deflate stream, crc32 + isize (8 bytes)
    if ((size_t)bsize + 1 < 12 + (size_t)xlen + 8) { r->error = "malformed BGZF block (BSIZE smaller than its header)"; r->eof = true; return false; }
    const size_t clen = (size_t)bsize + 1 - 12 - (size_t)xlen;   // compressed data + crc32 + isize
    std::vector<uint8_t> comp(clen);
    if (fread(comp.data(), 1, clen, r->f) != clen) { r->error = "truncated BGZF block"; r->eof = true; return false; }
    const uint32_t isize = comp[clen - 4] | (comp[clen - 3] << 8) | (comp[clen - 2] << 16) | ((uint32_t)comp[clen - 1] << 24);
    if (isize == 0) return true;   // the empty end-of-file block
    if (isize > 65536) { r->error = "malformed BGZF block (ISIZE above 64 KiB)"; r->eof = true; return false; }
    if (r->cur > 0 && r->cur == r->buf.size()) { r->buf.clear(); r->cur = 0; }
    const size_t old = r->buf.size();
    r->buf.resize(old + isize);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) { r->error = "inflateInit2 failed"; r->eof = true; return false; }
    zs.next_in = comp.data(); zs.avail_in = (uInt)(clen - 8);
    zs.next_out = r->buf.data() + old; zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != isize) { r->error = "BGZF inflate failed"; r->eof = true; return false; }
    return true;
}
// make n bytes available at r->cur
bool need(pb2_bam_reader* r, size_t n) {
    while (r->buf.size() - r->cur < n) {
        if (r->cur > (1u << 20)) { r->buf.erase(r->buf.begin(), r->buf.begin() + (long)r->cur); r->cur = 0; }
        if (!next_block(r)) return false;
    }
    return true;
}
int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }

struct Tags { bool has_xv = false, has_xw = false; int64_t xv = 0, xw = 0; bool has_xr = false, has_xd = false; std::string xr, xd; bool any = false;
              bool has_xn = false, bad_xn = false; std::string xn; bool malformed = false; };
Tags parse_tags(const uint8_t* p, size_t n) {
    Tags t;
    size_t i = 0;
    t.any = n > 0;
    while (i + 3 <= n) {
        const char a = (char)p[i], b = (char)p[i + 1], ty = (char)p[i + 2];
        i += 3;
        int64_t iv = 0;
        bool is_int = false;
        std::string sv;
        bool is_str = false;
        // every value is read inside the record: a tag that runs past its end makes the record malformed
        size_t fixed = 0;
        switch (ty) {
            case 'c': case 'C': case 'A': fixed = 1; break;
            case 's': case 'S': fixed = 2; break;
            case 'i': case 'I': case 'f': fixed = 4; break;
            case 'B': fixed = 5; break;
            default: break;
        }
        if (i + fixed > n) { t.malformed = true; return t; }
        if (ty == 'B') {
            const char st = (char)p[i];
            const int32_t cnt = rd_i32(p + i + 1);
            const size_t sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            if (cnt < 0 || (size_t)cnt > (n - i - 5) / sz) { t.malformed = true; return t; }
        }
        if (a == 'X' && b == 'N') {   // Read.GetAmpliconNameIfExists -> TagUtils.GetStringTag (BamCommon.cs:1182-1216): Z / H strings, A / C (upper-cased type)
            t.has_xn = true;          // one character, anything else throws
            const char up = (char)toupper((unsigned char)ty);
            if (up == 'A' || up == 'C') t.xn.assign(1, (char)p[i]);
            else if (up != 'Z' && up != 'H') t.bad_xn = true;
        }
        switch (ty) {
            case 'c': iv = (int8_t)p[i]; i += 1; is_int = true; break;
            case 'C': iv = p[i]; i += 1; is_int = true; break;
            case 's': { int16_t v; memcpy(&v, p + i, 2); iv = v; i += 2; is_int = true; break; }
            case 'S': { uint16_t v; memcpy(&v, p + i, 2); iv = v; i += 2; is_int = true; break; }
            case 'i': iv = rd_i32(p + i); i += 4; is_int = true; break;
            case 'I': iv = rd_u32(p + i); i += 4; is_int = true; break;
            case 'f': i += 4; break;
            case 'A': i += 1; break;
            case 'Z': case 'H': {
                size_t e = i;
                while (e < n && p[e] != 0) e++;
                if (e >= n) { t.malformed = true; return t; }   // no terminator inside the record
                sv.assign((const char*)p + i, e - i); i = e + 1; is_str = true; break;
            }
            case 'B': {
                const char st = (char)p[i];
                const int32_t cnt = rd_i32(p + i + 1);
                const int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                i += 5 + (size_t)sz * (size_t)std::max(cnt, 0);
                break;
            }
            default: return t;   // unknown type: stop
        }
        if (a == 'X' && b == 'V' && is_int) { t.has_xv = true; t.xv = iv; }
        else if (a == 'X' && b == 'W' && is_int) { t.has_xw = true; t.xw = iv; }
        else if (a == 'X' && b == 'R' && is_str) { t.has_xr = true; t.xr = sv; }
        else if (a == 'X' && b == 'D' && is_str) { t.has_xd = true; t.xd = sv; }
        else if (a == 'X' && b == 'N' && is_str) t.xn = sv;
    }
    return t;
}
}  // namespace

static int bam_open_impl(const char* path, pb2_bam_reader** out);
extern "C" int pb2_bam_open(const char* path, pb2_bam_reader** out) {
    try { return bam_open_impl(path, out); }
    catch (const std::exception&) { if (out) *out = nullptr; return PB2_ERR_NOMEM; }
}
static int bam_open_impl(const char* path, pb2_bam_reader** out) {
    if (!path || !out) return PB2_ERR_ARG;
    *out = nullptr;
    pb2_bam_reader* r = new pb2_bam_reader();
    r->f = fopen(path, "rb");
    if (!r->f) { delete r; return PB2_ERR_ARG; }
    auto bad = [&](const char* m) { r->error = m; fclose(r->f); delete r; return PB2_ERR_ARG; };
    if (!need(r, 12) || memcmp(r->buf.data() + r->cur, "BAM\1", 4) != 0) return bad("not a BAM file");
    const int32_t l_text = rd_i32(r->buf.data() + r->cur + 4);
    if (l_text < 0 || !need(r, 12 + (size_t)l_text)) return bad("truncated BAM header");
    r->header_text.assign((const char*)r->buf.data() + r->cur + 8, (size_t)l_text);
    const int32_t n_ref = rd_i32(r->buf.data() + r->cur + 8 + l_text);
    r->cur += 12 + (size_t)l_text;
    for (int32_t i = 0; i < n_ref; i++) {
        if (!need(r, 4)) return bad("truncated BAM reference list");
        const int32_t l_name = rd_i32(r->buf.data() + r->cur);
        if (l_name < 1 || !need(r, 8 + (size_t)l_name)) return bad("truncated BAM reference list");
        r->ref_names.emplace_back((const char*)r->buf.data() + r->cur + 4, (size_t)l_name - 1);
        r->ref_lengths.push_back(rd_i32(r->buf.data() + r->cur + 4 + l_name));
        r->cur += 8 + (size_t)l_name;
    }
    for (auto& s : r->ref_names) r->ref_name_ptrs.push_back(s.c_str());
    // BamFileAlignmentExtractor.CheckBamHeaderIfBamHasBeenStitched / CheckIfBamHasBeenCollapsed (:111-153)
    size_t a = 0;
    while (a <= r->header_text.size()) {
        size_t e = r->header_text.find('\n', a);
        if (e == std::string::npos) e = r->header_text.size();
        const std::string line = r->header_text.substr(a, e - a);
        if (line.size() >= 3 && line.compare(0, 3, "@PG") == 0) {
            const std::string l = lower(line);
            if (l.find("stitcher") != std::string::npos && l.find("pisces") != std::string::npos) r->is_stitched = 1;
            if (l.find("pn:reco") != std::string::npos) r->is_collapsed = 1;
        }
        a = e + 1;
    }
    *out = r;
    return PB2_OK;
}
extern "C" void pb2_bam_close(pb2_bam_reader* r) { if (r) { if (r->f) fclose(r->f); delete r; } }
extern "C" const char* pb2_bam_last_error(pb2_bam_reader* r) { return r ? r->error.c_str() : "pb2_bam_open failed"; }
extern "C" int pb2_bam_header(pb2_bam_reader* r, int32_t* n_refs, const char* const** names, const int32_t** lengths, int32_t* is_stitched, int32_t* is_collapsed) {
    if (!r) return PB2_ERR_ARG;
    if (n_refs) *n_refs = (int32_t)r->ref_names.size();
    if (names) *names = r->ref_name_ptrs.data();
    if (lengths) *lengths = r->ref_lengths.data();
    if (is_stitched) *is_stitched = r->is_stitched;
    if (is_collapsed) *is_collapsed = r->is_collapsed;
    return PB2_OK;
}

static int bam_next_batch_impl(pb2_bam_reader* r, const pb2_bam_filter* flt, int32_t max_reads, pb2_read_batch* batch, int32_t* ref_id_out, int64_t* n_skipped);
// nothing throws across the C boundary: allocation failures and the like become error codes
extern "C" int pb2_bam_next_batch(pb2_bam_reader* r, const pb2_bam_filter* flt, int32_t max_reads, pb2_read_batch* batch, int32_t* ref_id_out, int64_t* n_skipped) {
    try { return bam_next_batch_impl(r, flt, max_reads, batch, ref_id_out, n_skipped); }
    catch (const std::bad_alloc&) { if (r) r->error = "out of memory while decoding the BAM file"; return PB2_ERR_NOMEM; }
    catch (const std::exception& e) { if (r) r->error = std::string("malformed BAM file: ") + e.what(); return PB2_ERR_ARG; }
}
static int bam_next_batch_impl(pb2_bam_reader* r, const pb2_bam_filter* flt, int32_t max_reads, pb2_read_batch* batch, int32_t* ref_id_out, int64_t* n_skipped) {
    if (!r || !batch || max_reads <= 0) return bfail(r, "pb2_bam_next_batch: bad argument");
    pb2_bam_filter f;
    f.min_map_quality = 1; f.remove_duplicates = 1; f.only_proper_pairs = 0;   // BamFilterParameters.cs:7-11
    if (flt) f = *flt;
    r->pos0.clear(); r->flag.clear(); r->cigar.clear(); r->bases.clear(); r->quals.clear(); r->base_dirs.clear(); r->collapsed.clear(); r->amplicon_id.clear();
    r->cigar_off.assign(1, 0); r->seq_off.assign(1, 0);
    bool any_dirs = false, any_coll = false;   // (kept for diagnostics)
    int64_t skipped = 0;
    int32_t batch_ref = -2;
    static const char kSeq[] = "=ACMGRSVTWYHKDBN";
    std::vector<uint8_t> rec;
    while ((int32_t)r->pos0.size() < max_reads) {
        if (r->have_pending) { rec.swap(r->pending); r->have_pending = false; }
        else {
            if (!need(r, 4)) break;
            const int32_t bs = rd_i32(r->buf.data() + r->cur);
            if (bs < 32 || !need(r, 4 + (size_t)bs)) { if (r->error.empty()) r->error = "truncated BAM record"; return PB2_ERR_ARG; }
            rec.assign(r->buf.begin() + (long)r->cur + 4, r->buf.begin() + (long)r->cur + 4 + bs);
            r->cur += 4 + (size_t)bs;
        }
        const uint8_t* p = rec.data();
        const int32_t ref_id = rd_i32(p), pos = rd_i32(p + 4);
        const uint32_t bin_mq_nl = rd_u32(p + 8), flag_nc = rd_u32(p + 12);
        const int32_t l_seq = rd_i32(p + 16);
        const int l_name = bin_mq_nl & 0xff, mapq = (bin_mq_nl >> 8) & 0xff;
        const uint32_t flag = flag_nc >> 16, n_cig = flag_nc & 0xffff;
        const size_t o_cig = 32 + (size_t)l_name, o_seq = o_cig + 4 * (size_t)n_cig, o_qual = o_seq + ((size_t)l_seq + 1) / 2, o_tags = o_qual + (size_t)l_seq;
        if (l_seq < 0 || o_tags > rec.size()) return bfail(r, "malformed BAM record");
        // AlignmentSource.ShouldSkipRead (:84-92): !IsMapped, !IsPrimaryAlignment, (OnlyUseProperPairs && !IsProperPair), (RemoveDuplicates && IsPcrDuplicate),
        // MapQuality < MinimumMapQuality, no CIGAR
        const bool skip = (flag & 0x4) || (flag & 0x100) || (f.only_proper_pairs && !(flag & 0x2)) || (f.remove_duplicates && (flag & 0x400)) || mapq < f.min_map_quality ||
                          n_cig == 0 || ref_id < 0;
        if (skip) { skipped++; continue; }
        if (batch_ref == -2) batch_ref = ref_id;
        else if (ref_id != batch_ref) { r->pending.swap(rec); r->have_pending = true; break; }   // the next chromosome starts: its reads go to the next batch
        const Tags t = parse_tags(p + o_tags, rec.size() - o_tags);
        if (t.malformed) return bfail(r, "malformed BAM record (a tag runs past the end of the record)");
        if (t.bad_xn) return bfail(r, "Found an unexpected string BAM tag data type while looking for a tag (XN)");
        int32_t amp = -1;
        if (t.has_xn) {
            for (size_t k = 0; k < r->amplicon_names.size() && amp < 0; k++) if (r->amplicon_names[k] == t.xn) amp = (int32_t)k;
            if (amp < 0) { amp = (int32_t)r->amplicon_names.size(); r->amplicon_names.push_back(t.xn); }
        }
        r->amplicon_id.push_back(amp);
        r->pos0.push_back(pos);
        r->flag.push_back((uint16_t)flag);
        for (uint32_t k = 0; k < n_cig; k++) r->cigar.push_back(rd_u32(p + o_cig + 4 * k));
        r->cigar_off.push_back((int64_t)r->cigar.size());
        for (int32_t i = 0; i < l_seq; i++) {
            const uint8_t b = p[o_seq + ((size_t)i >> 1)];
            r->bases.push_back((uint8_t)kSeq[(i & 1) ? (b & 15) : (b >> 4)]);
            r->quals.push_back(p[o_qual + (size_t)i]);
        }
        // Read.SequencedBaseDirectionMap: the XD runs ("12F30S8R") walked along the expanded CIGAR, kept for the operations that span the read
        const bool reverse = (flag & 0x10) != 0;
        const size_t d0 = r->base_dirs.size();
        r->base_dirs.resize(d0 + (size_t)l_seq, (uint8_t)(reverse ? 1 : 0));
        if (t.has_xd && !t.xd.empty()) {
            std::vector<uint8_t> expanded;
            size_t cigar_units = 0;   // the direction string is walked along the expanded CIGAR: runs beyond its length are never looked at
            for (uint32_t k = 0; k < n_cig; k++) cigar_units += rd_u32(p + o_cig + 4 * k) >> 4;
            int64_t num = 0;
            for (char ch : t.xd) {
                if (ch >= '0' && ch <= '9') num = std::min<int64_t>(num * 10 + (ch - '0'), (int64_t)1 << 40);
                else {
                    const uint8_t dv = ch == 'F' ? 0 : ch == 'R' ? 1 : 2;
                    expanded.insert(expanded.end(), (size_t)std::min<int64_t>(num, (int64_t)(cigar_units - std::min(cigar_units, expanded.size()))), dv);
                    num = 0;
                }
            }
            size_t ci = 0, si = 0;
            for (uint32_t k = 0; k < n_cig; k++) {
                const uint32_t c = rd_u32(p + o_cig + 4 * k);
                const int op = c & 15;
                const bool read_span = op == 0 || op == 1 || op == 4 || op == 7 || op == 8;
                for (uint32_t j = 0; j < (c >> 4); j++, ci++)
                    if (read_span && si < (size_t)l_seq) { r->base_dirs[d0 + si] = ci < expanded.size() ? expanded[ci] : (uint8_t)(reverse ? 1 : 0); si++; }
            }
            any_dirs = true;
        }
        // IsCollapsedRead (XV or XW present), IsDuplex (both non-zero), ReadPairDirection (XR, else from the flags of a proper pair)
        std::string xr = t.has_xr ? t.xr : std::string();
        if (!t.has_xr && (flag & 0x2)) {
            const char dir = reverse ? 'R' : 'F', mate = reverse ? 'F' : 'R';
            xr = (flag & 0x40) ? std::string{dir, mate} : std::string{mate, dir};
        }
        const bool collapsed_read = t.has_xv || t.has_xw;
        const bool duplex = t.has_xv && t.xv != 0 && t.has_xw && t.xw != 0;
        r->collapsed.push_back((uint8_t)((collapsed_read ? 1 : 0) | (duplex ? 2 : 0) | ((xr == "FR" ? 1 : xr == "RF" ? 2 : 0) << 2)));
        any_coll |= collapsed_read;
        r->seq_off.push_back((int64_t)r->bases.size());
    }
    if (!r->error.empty()) return PB2_ERR_ARG;
    memset(batch, 0, sizeof(*batch));
    batch->n_reads = (int32_t)r->pos0.size();
    batch->pos0 = r->pos0.data(); batch->flag = r->flag.data(); batch->cigar_off = r->cigar_off.data(); batch->cigar = r->cigar.data();
    batch->seq_off = r->seq_off.data(); batch->bases = r->bases.data(); batch->quals = r->quals.data();
    // always handed out (both hold the values the flags imply where a read carries no XD / XV / XW tag): a stitched or collapsed BAM streamed in small
    // batches mixes tagged and untagged reads, and a batch must not change shape with its content
    r->batch_any_dirs = any_dirs; r->batch_any_coll = any_coll;
    batch->base_dirs = r->base_dirs.data();
    batch->collapsed = r->collapsed.data();
    // amplicon name ids once the file has shown an XN tag (a file without the tag never asks the caller to track amplicons)
    batch->amplicon = r->amplicon_names.empty() ? nullptr : r->amplicon_id.data();
    if (ref_id_out) *ref_id_out = batch_ref == -2 ? -1 : batch_ref;
    if (n_skipped) *n_skipped = skipped;
    return PB2_OK;
}

// The same batch in the packed form of pb2_push_reads_packed: one byte per base + the exception list, compact offsets (operation counts) when every read
// has at most 255 CIGAR operations, per-base directions only when a read of the batch carried an XD tag and collapsed summaries only when one carried
// XV / XW (pb2_push_reads_packed gives the other reads their flag-derived defaults): 1.1 bytes per base cross the PCIe link instead of 3.
extern "C" int pb2_bam_next_batch_packed(pb2_bam_reader* r, const pb2_bam_filter* flt, int32_t max_reads, pb2_packed_read_batch* out, int32_t* ref_id_out, int64_t* n_skipped) {
    if (!r || !out) return PB2_ERR_ARG;
    try {
        pb2_read_batch b;
        const int rc = bam_next_batch_impl(r, flt, max_reads, &b, ref_id_out, n_skipped);
        if (rc != PB2_OK) return rc;
        memset(out, 0, sizeof(*out));
        out->n_reads = b.n_reads;
        if (b.n_reads == 0) return PB2_OK;
        const int64_t n_seq = (int64_t)r->bases.size();
        r->packed_seq.resize((size_t)n_seq);
        int64_t cap = std::max<int64_t>(1024, n_seq / 64);
        for (;;) {
            r->exc_index.resize((size_t)cap); r->exc_base.resize((size_t)cap); r->exc_qual.resize((size_t)cap);
            const int64_t ne = pb2_pack_reads(r->bases.data(), r->quals.data(), n_seq, r->packed_seq.data(), r->exc_index.data(), r->exc_base.data(), r->exc_qual.data(), cap);
            if (ne < 0) return bfail(r, "pb2_pack_reads failed");
            if (ne <= cap) { out->n_exceptions = ne; break; }
            cap = ne;
        }
        bool compact = true;
        r->cigar_ops.resize((size_t)b.n_reads);
        for (int32_t i = 0; i < b.n_reads; i++) {
            const int64_t ops = r->cigar_off[(size_t)i + 1] - r->cigar_off[(size_t)i];
            if (ops > 255) { compact = false; break; }
            r->cigar_ops[(size_t)i] = (uint8_t)ops;
        }
        out->pos0 = b.pos0; out->flag = b.flag; out->cigar = b.cigar; out->seq = r->packed_seq.data();
        out->exc_index = r->exc_index.data(); out->exc_base = r->exc_base.data(); out->exc_qual = r->exc_qual.data();
        if (compact) { out->cigar_ops = r->cigar_ops.data(); out->n_cigar_total = (int64_t)r->cigar.size(); out->n_seq_total = n_seq; }
        else { out->cigar_off = b.cigar_off; out->seq_off = b.seq_off; }
        out->base_dirs = r->batch_any_dirs ? b.base_dirs : nullptr;
        out->collapsed = r->batch_any_coll ? b.collapsed : nullptr;
        out->amplicon = b.amplicon;
        return PB2_OK;
    }
    catch (const std::bad_alloc&) { r->error = "out of memory while decoding the BAM file"; return PB2_ERR_NOMEM; }
    catch (const std::exception& e) { r->error = std::string("malformed BAM file: ") + e.what(); return PB2_ERR_ARG; }
}

// Amplicon names of the reads (the XN tag, Read.GetAmpliconNameIfExists, src/lib/Pisces.Domain/Models/Read.cs:479-486): ids of the reads of the batch
// handed out last, and the file's dictionary so far (first-seen order over the kept reads). Input of the amplicon-bias row.
extern "C" int pb2_bam_batch_amplicons(pb2_bam_reader* r, const int32_t** amplicon_id, int32_t* n_reads) {
    if (!r || !amplicon_id || !n_reads) return PB2_ERR_ARG;
    *amplicon_id = r->amplicon_id.data();
    *n_reads = (int32_t)r->amplicon_id.size();
    return PB2_OK;
}
extern "C" int pb2_bam_amplicon_names(pb2_bam_reader* r, int32_t* n, const char* const** names) {
    if (!r || !n || !names) return PB2_ERR_ARG;
    r->amplicon_name_ptrs.clear();
    for (auto& s : r->amplicon_names) r->amplicon_name_ptrs.push_back(s.c_str());
    *n = (int32_t)r->amplicon_names.size();
    *names = r->amplicon_name_ptrs.data();
    return PB2_OK;
}
