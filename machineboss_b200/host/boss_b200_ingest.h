// boss_b200_ingest.h -- batch ingest at scale: sequence files straight to the packed token
// arrays the device takes (mb_batch_create), without a std::string per residue on the way.
//
// The reference reads a SeqPairList as a JSON tree, validates every pair and every sequence against a schema that it
// re-parses each time, and keeps each residue as a std::string in a vector (src/seqpair.cpp:8-38,245-249,
// src/seqpair.h:37-44, src/schema.cpp:74-92); the matrices then tokenise per use (dpmatrix.defs.h:6-7).  At 10^5 - 10^6
// pairs that is where the time goes once the DP runs on a GPU.  Here:
//   * packedFromSeqPairListJson: one streaming pass over the reference's JSON list format
//         [ { "input": { "name": .., "sequence": [sym, ..] }, "output": { .. } [, "meta": ..] }, .. ]
//     tokenising each symbol as it is read (pairs that carry an "alignment" need envelopes: the general reader takes those);
//   * packedFromFasta: two FASTA files, record k of the first paired with record k of the second, one residue per character.
// Unknown symbols fail like Tokenizer::tokenize (eval.h:33-37).
#ifndef MB_HOST_BOSS_B200_INGEST_H
#define MB_HOST_BOSS_B200_INGEST_H

#include <cstring>

#include "boss_b200.h"

namespace MachineBoss {

struct PackedPairs {
  vector<uint8_t> x, y;                 // tokens, 1-based, concatenated
  vector<int64_t> xOff, yOff;           // [nPairs + 1]
  vector<string> xName, yName;
  PackedPairs() : xOff (1, 0), yOff (1, 0) {}
  size_t size() const { return xOff.size() - 1; }
  size_t residues() const { return x.size() + y.size(); }
};

namespace ingest {
  // symbol -> token: a 256-entry table when every symbol of the alphabet is one character (DNA, protein), the tokenizer's map otherwise
  template<class Tok>
  struct FastTokens {
    const Tok& tok;
    uint8_t byChar[256];
    bool singleChar = true;
    explicit FastTokens (const Tok& t) : tok (t) {
      memset (byChar, 0, sizeof byChar);
      for (size_t k = 1; k < t.tok2sym.size(); ++k) {
        if (t.tok2sym[k].size() != 1) { singleChar = false; continue; }
        byChar[(unsigned char) t.tok2sym[k][0]] = (uint8_t) k;
      }
    }
    uint8_t of (const char* sym, size_t len) const {
      if (len == 1 && singleChar) { const uint8_t v = byChar[(unsigned char) sym[0]]; if (v) return v; }
      else { const auto it = tok.sym2tok.find (string (sym, len)); if (it != tok.sym2tok.end() && it->second > 0) return (uint8_t) it->second; }
      std::ostringstream err;
      err << "Can't tokenize symbol " << string (sym, len) << " using this alphabet:";
      for (const auto& s: tok.tok2sym) err << ' ' << s;
      throw runtime_error (err.str());
    }
  };

  struct Scanner {      // just enough JSON for the list format; strings without escapes are taken in place
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool eat (char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    void need (char c) { if (!eat (c)) throw runtime_error (string ("SeqPairList JSON: expected '") + c + "'"); }
    // a string: [s, s + n) points into the text when it has no escape, into `scratch` otherwise
    void str (const char*& s, size_t& n, string& scratch) {
      need ('"');
      const char* b = p;
      while (p < end && *p != '"' && *p != '\\') ++p;
      if (p < end && *p == '"') { s = b; n = (size_t) (p - b); ++p; return; }
      scratch.assign (b, p);
      while (p < end && *p != '"') {
        if (*p == '\\' && p + 1 < end) {
          ++p;
          switch (*p) { case 'n': scratch += '\n'; break; case 't': scratch += '\t'; break; case 'r': scratch += '\r'; break; case 'b': scratch += '\b'; break; case 'f': scratch += '\f'; break;
                        case 'u': scratch += '?'; p += std::min<ptrdiff_t> (4, end - p - 1); break; default: scratch += *p; }
          ++p;
        } else scratch += *p++;
      }
      if (p >= end) throw runtime_error ("SeqPairList JSON: unterminated string");
      ++p;
      s = scratch.data(); n = scratch.size();
    }
    void skipValue() {
      ws();
      if (p >= end) throw runtime_error ("SeqPairList JSON: unexpected end");
      if (*p == '"') { const char* s; size_t n; string scratch; str (s, n, scratch); return; }
      if (*p == '{' || *p == '[') {
        const char open = *p, close = open == '{' ? '}' : ']';
        ++p;
        if (eat (close)) return;
        do { if (open == '{') { const char* s; size_t n; string scratch; str (s, n, scratch); need (':'); } skipValue(); } while (eat (','));
        need (close);
        return;
      }
      while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') ++p;      // number, true, false, null
    }
  };
}  // namespace ingest

inline PackedPairs packedFromSeqPairListJson (const string& text, const EvaluatedMachine& m) {
  PackedPairs pp;
  const ingest::FastTokens<InputTokenizer> inTok (m.inputTokenizer);
  const ingest::FastTokens<OutputTokenizer> outTok (m.outputTokenizer);
  ingest::Scanner sc { text.data(), text.data() + text.size() };
  string scratch;
  const char* s; size_t n;
  sc.need ('[');
  if (!sc.eat (']')) {
    do {
      sc.need ('{');
      string xName, yName;
      if (!sc.eat ('}')) {
        do {
          sc.str (s, n, scratch);
          const string key (s, n);
          sc.need (':');
          if (key == "input" || key == "output") {
            const bool isIn = key == "input";
            sc.need ('{');
            if (!sc.eat ('}')) {
              do {
                sc.str (s, n, scratch);
                const bool isName = n == 4 && !memcmp (s, "name", 4), isSeq = n == 8 && !memcmp (s, "sequence", 8);
                sc.need (':');
                if (isName) { sc.str (s, n, scratch); (isIn ? xName : yName).assign (s, n); }
                else if (isSeq) {
                  sc.need ('[');
                  if (!sc.eat (']')) {
                    do { sc.str (s, n, scratch); if (isIn) pp.x.push_back (inTok.of (s, n)); else pp.y.push_back (outTok.of (s, n)); } while (sc.eat (','));
                    sc.need (']');
                  }
                } else sc.skipValue();
              } while (sc.eat (','));
              sc.need ('}');
            }
          } else if (key == "alignment") throw runtime_error ("SeqPairList JSON: a pair carries an alignment (it needs an envelope): use the general reader");
          else sc.skipValue();
        } while (sc.eat (','));
        sc.need ('}');
      }
      pp.xOff.push_back ((int64_t) pp.x.size());
      pp.yOff.push_back ((int64_t) pp.y.size());
      pp.xName.push_back (xName);
      pp.yName.push_back (yName);
    } while (sc.eat (','));
    sc.need (']');
  }
  return pp;
}

namespace ingest {
  template<class Tok>
  inline void fasta (const string& text, const FastTokens<Tok>& tok, vector<uint8_t>& seq, vector<int64_t>& off, vector<string>& names) {
    const char* p = text.data(); const char* end = p + text.size();
    seq.resize (text.size());      // at most one token per byte of the file: written through a pointer, trimmed at the end
    uint8_t* out = seq.data();
    bool open = false;
    while (p < end) {
      const char* eol = (const char*) memchr (p, '\n', (size_t) (end - p));
      if (!eol) eol = end;
      if (p < eol && *p == '>') {
        if (open) off.push_back ((int64_t) (out - seq.data()));
        open = true;
        const char* e = p + 1;
        while (e < eol && *e != ' ' && *e != '\t' && *e != '\r') ++e;
        names.push_back (string (p + 1, e));
      } else if (open) {
        if (tok.singleChar) {
          for (const char* q = p; q < eol; ++q) {
            const uint8_t v = tok.byChar[(unsigned char) *q];
            if (v) *out++ = v;
            else if (*q != ' ' && *q != '\t' && *q != '\r') tok.of (q, 1);      // throws: not in the alphabet
          }
        } else for (const char* q = p; q < eol; ++q) if (*q != ' ' && *q != '\t' && *q != '\r') *out++ = tok.of (q, 1);
      }
      p = eol + 1;
    }
    if (open) off.push_back ((int64_t) (out - seq.data()));
    seq.resize ((size_t) (out - seq.data()));
  }
}

inline PackedPairs packedFromFasta (const string& inputText, const string& outputText, const EvaluatedMachine& m) {
  PackedPairs pp;
  ingest::fasta (inputText, ingest::FastTokens<InputTokenizer> (m.inputTokenizer), pp.x, pp.xOff, pp.xName);
  ingest::fasta (outputText, ingest::FastTokens<OutputTokenizer> (m.outputTokenizer), pp.y, pp.yOff, pp.yName);
  if (pp.xOff.size() != pp.yOff.size()) throw runtime_error ("paired FASTA files hold different numbers of sequences");
  return pp;
}

// the packed list on every GPU of the box (as ListBatch, without SeqPair objects): Forward log-likelihoods / Viterbi scores
inline vector<double> packedScores (const EvaluatedMachine& m, const PackedPairs& pp, bool viterbi) {
  const int64_t n = (int64_t) pp.size();
  vector<double> sc ((size_t) n);
  if (!n) return sc;
  vector<uint8_t> x = pp.x, y = pp.y;
  x.push_back (0); y.push_back (0);
  mb_group* g = n > 1 ? hostGroup() : nullptr;
  if (g) {
    mb_gbatch* gb = nullptr;
    mbCheck (mb_group_batch_create (g, &gb, n, x.data(), pp.xOff.data(), y.data(), pp.yOff.data()));
    const int rc = viterbi ? mb_group_viterbi (m.groupHandle (g), gb, sc.data(), nullptr) : mb_group_forward (m.groupHandle (g), gb, sc.data());
    mb_group_batch_destroy (gb);
    mbCheck (rc);
  } else {
    mb_batch* b = nullptr;
    mbCheck (mb_batch_create (&b, n, x.data(), pp.xOff.data(), y.data(), pp.yOff.data()));
    const int rc = viterbi ? mb_viterbi (m.handle(), b, sc.data(), nullptr) : mb_forward (m.handle(), b, sc.data());
    mb_batch_destroy (b);
    mbCheck (rc);
  }
  return sc;
}

}  // namespace MachineBoss

#endif
