// CPU ORACLE (test infrastructure only) -- canonical Huffman coder over the 260-symbol integer alphabet.
// Restates C/compress/canonicalHuffman/{CanonicalHuffman,TreeBuilder,PackageMerge,LengthEncoder,
// HuffmanCodeBits,CanonHuffTreeDecoder}.java  (C/ = /root/reference/core/src/main/java/org/gridfour/).
#include "g4oracle.h"
#include <algorithm>

namespace g4o {

namespace {
constexpr int N_SYMBOLS_TOTAL = 260;      // CanonicalHuffman.java:74
constexpr int I_NULL_DATA_CODE = 256;     // :77
constexpr int I_ESCAPE_1BYTE = 257;       // :78
constexpr int I_ESCAPE_2BITS = 258;       // :79
constexpr int I_END_OF_TEXT = 259;        // :80
constexpr int MAX_STANDARD_SYMBOL = 15;   // LengthEncoder.java:50
constexpr int REPEAT_PREV_2BITS = 16;     // :56
constexpr int REPEAT_ZERO_3BITS = 17;     // :61
constexpr int REPEAT_ZERO_7BITS = 18;     // :66
constexpr int SYMBOL_SET_SIZE = 19;       // :72

struct Code {
  int nBits = 0;
  uint64_t bits = 0;  // canonical code value; emitted MSB-first (HuffmanCodeBits.java:67-73)
};

struct TNode {
  bool isLeaf = false;
  int symbol = -1;
  int count = 0;
  int next = -1, left = -1, right = -1;
  int nBitsInCode = 0;
};

// TreeBuilder.establishCodeLengths (TreeBuilder.java:200-274): leaf depth by pre-order walk.
int establish_code_lengths(std::vector<TNode>& nodes, int root, int nSymbols) {
  int maxLen = 0;
  std::vector<int> path(size_t(nSymbols) + 2), branch(size_t(nSymbols) + 2);
  path[0] = root; branch[0] = 0;
  int depth = 1;
  while (depth > 0) {
    int index = depth - 1;
    int p = path[index];
    switch (branch[index]) {
      case 0:
        if (nodes[p].isLeaf) {
          depth--;
          nodes[p].nBitsInCode = depth;
          if (depth > maxLen) maxLen = depth;
        } else {
          branch[index] = 1; branch[depth] = 0; path[depth] = nodes[p].left; depth++;
        }
        break;
      case 1:
        branch[index] = 2; branch[depth] = 0; path[depth] = nodes[p].right; depth++;
        break;
      default:
        branch[index] = 0; depth--;
        break;
    }
  }
  return maxLen;
}
}  // namespace

// PackageMerge.merge (PackageMerge.java:91-175).  `sortedCounts` are the counts of the sortNodes
// array in TreeBuilder order (count asc, symbol desc); entry "symbol" == index into that array, so
// the (count, index) sort at :106-112 is the identity permutation on it.
void package_merge(int maxLen, const int* sortedCounts, int n, int* nBitsOut) {
  struct Entry { int symbol; int count; int nBits; };
  std::vector<Entry> baseStore;
  for (int i = 0; i < n; i++) if (sortedCounts[i] > 0) baseStore.push_back({i, sortedCounts[i], 0});
  std::stable_sort(baseStore.begin(), baseStore.end(), [](const Entry& a, const Entry& b) {
    if (a.count != b.count) return a.count < b.count;
    return a.symbol < b.symbol;
  });
  const int B = int(baseStore.size());
  // entries[level] holds pointers: >=0 -> index into baseStore, -1 -> package, -2 -> null slot
  struct Ref { int base; int count; };
  std::vector<std::vector<Ref>> entries(maxLen);
  entries[0].resize(B);
  for (int i = 0; i < B; i++) entries[0][i] = {i, baseStore[i].count};
  for (int d = 1; d < maxLen; d++) {
    const std::vector<Ref>& ix = entries[d - 1];
    int nPair = int(ix.size()) / 2;
    std::vector<int> pair(nPair);
    for (int i = 0; i < nPair; i++) pair[i] = ix[2 * i].count + ix[2 * i + 1].count;
    std::vector<Ref> m(size_t(B) + nPair, Ref{-2, 0});
    int k = 0, iBase = 0;
    for (int iPair = 0; iPair < nPair; iPair++) {
      while (iBase < B) {
        if (baseStore[iBase].count <= pair[iPair]) { m[k++] = {iBase, baseStore[iBase].count}; iBase++; }
        else break;
      }
      m[k++] = {-1, pair[iPair]};
    }
    if (baseStore[B - 1].count > pair[nPair - 1]) m[m.size() - 1] = {B - 1, baseStore[B - 1].count};  // :145-147
    entries[d] = m;
  }
  int nn = B * 2 - 2;
  for (int e = maxLen - 1; e >= 0; e--) {
    int nMerged = 0;
    const std::vector<Ref>& ix = entries[e];
    for (int i = 0; i < nn; i++) {
      if (ix.at(i).base == -2) throw std::runtime_error("package_merge: null entry (Java NPE)");
      if (ix[i].base == -1) nMerged++;
      else baseStore[ix[i].base].nBits++;
    }
    nn = nMerged * 2;
  }
  for (int i = 0; i < n; i++) nBitsOut[i] = 0;
  for (int i = 0; i < B; i++) nBitsOut[baseStore[i].symbol] = baseStore[i].nBits;
}

// TreeBuilder.buildTree (TreeBuilder.java:75-188) reduced to what defines the stream: the code length of
// every symbol.  counts[i] is the count of symbol i.
void canon_tree_lengths(const int* counts, int nSymbols, int* lengths, bool* limited) {
  std::vector<TNode> nodes(nSymbols);
  std::vector<int> sortNodes;
  for (int i = 0; i < nSymbols; i++) {
    nodes[i].isLeaf = true; nodes[i].symbol = i; nodes[i].count = counts[i];
    lengths[i] = 0;
    if (counts[i] > 0) sortNodes.push_back(i);
  }
  if (limited) *limited = false;
  // count ascending, symbol DESCENDING (TreeBuilder.java:100-130)
  std::sort(sortNodes.begin(), sortNodes.end(), [&](int a, int b) {
    if (nodes[a].count != nodes[b].count) return nodes[a].count < nodes[b].count;
    return nodes[a].symbol > nodes[b].symbol;
  });
  const int k = int(sortNodes.size());
  if (k < 2) throw std::invalid_argument("canon tree needs >= 2 symbols");  // Java: NullPointerException
  for (int i = 0; i < k - 1; i++) nodes[sortNodes[i]].next = sortNodes[i + 1];
  int first = sortNodes[0];
  int root = -1;
  while (true) {  // TreeBuilder.java:139-169 (same insertion rule as the legacy encoder)
    int left = first;
    int right = nodes[first].next;
    first = nodes[right].next;
    nodes[left].next = -1; nodes[right].next = -1;
    TNode br; br.left = left; br.right = right; br.count = nodes[left].count + nodes[right].count;
    int b = int(nodes.size());
    nodes.push_back(br);
    if (first < 0) { root = b; break; }
    if (nodes[first].count >= nodes[b].count) { nodes[b].next = first; first = b; }
    else {
      int node = nodes[first].next, prior = first;
      while (node >= 0 && nodes[node].count < nodes[b].count) { prior = node; node = nodes[node].next; }
      nodes[prior].next = b;
      if (node >= 0) nodes[b].next = node;
    }
  }
  int maxLen = establish_code_lengths(nodes, root, k);
  if (maxLen > MAX_STANDARD_SYMBOL) {  // TreeBuilder.java:173-178
    if (limited) *limited = true;
    std::vector<int> sc(k), nb(k);
    for (int i = 0; i < k; i++) sc[i] = nodes[sortNodes[i]].count;
    package_merge(MAX_STANDARD_SYMBOL, sc.data(), k, nb.data());
    for (int i = 0; i < k; i++) nodes[sortNodes[i]].nBitsInCode = nb[i];
  }
  for (int i = 0; i < nSymbols; i++) lengths[i] = counts[i] > 0 ? nodes[i].nBitsInCode : 0;
}

namespace {
// TreeBuilder.populateCanonicalCodes (TreeBuilder.java:283-301) + HuffmanCodeBits.java:47-65
void canonical_codes(const int* lengths, int nSymbols, std::vector<Code>& codes) {
  codes.assign(nSymbols, Code());
  std::vector<int> order;
  for (int i = 0; i < nSymbols; i++) if (lengths[i] > 0) order.push_back(i);
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (lengths[a] != lengths[b]) return lengths[a] < lengths[b];
    return a < b;
  });
  uint64_t bits = 0;
  int prevLen = 0;
  for (size_t i = 0; i < order.size(); i++) {
    int len = lengths[order[i]];
    if (i == 0) { bits = 0; prevLen = len; }
    else {
      bits = bits + 1;
      if (len > prevLen) { bits <<= (len - prevLen); prevLen = len; }
    }
    codes[order[i]].bits = bits;
    codes[order[i]].nBits = prevLen;
  }
}

inline void write_code(BitOut& out, const Code& c) {  // TreeBuilder.writeOneSymbol :303-318
  for (int i = c.nBits - 1; i >= 0; i--) out.appendBit(int((c.bits >> i) & 1));
}
}  // namespace

// LengthEncoder.encodeLengths (LengthEncoder.java:86-167).  Returns nCodes.
int length_encode(int n, const int* codeLen, int* codes, int* runs) {
  int prior = -1;
  int i;
  int nc = 0;
  for (int k = 0; k < n; k++) { codes[k] = 0; runs[k] = 0; }
  for (int ic = 0; ic < n; ic++) {
    if (codeLen[ic] > MAX_STANDARD_SYMBOL) throw std::invalid_argument("Invalid code length");
    if (codeLen[ic] == 0) {
      prior = 0;
      for (i = ic + 1; i < n; i++) if (codeLen[i] != 0) break;
      int nZero = i - ic;
      if (nZero == 1) { codes[nc++] = 0; }
      else if (nZero == 2) { codes[nc++] = 0; codes[nc++] = 0; ic++; }
      else if (nZero <= 10) { codes[nc] = REPEAT_ZERO_3BITS; runs[nc] = nZero - 3; nc++; ic = i - 1; }
      else {
        if (nZero > 138) nZero = 138;
        codes[nc] = REPEAT_ZERO_7BITS; runs[nc] = nZero - 11; nc++;
        ic += (nZero - 1);
      }
    } else {
      if (codeLen[ic] == prior) {
        for (i = ic + 1; i < n; i++) if (codeLen[i] != prior) break;
        int nPrior = i - ic;
        switch (nPrior) {
          case 1: codes[nc++] = prior; break;
          case 2: codes[nc++] = prior; codes[nc++] = prior; ic = i - 1; break;
          default:
            if (nPrior > 6) nPrior = 6;
            codes[nc] = REPEAT_PREV_2BITS; runs[nc] = nPrior - 3; nc++;
            ic += (nPrior - 1);
            break;
        }
      } else {
        prior = codeLen[ic];
        codes[nc++] = prior;
      }
    }
  }
  return nc;
}

namespace {
inline void write_run_extra(BitOut& out, int code, int run) {
  switch (code) {
    case REPEAT_PREV_2BITS: out.appendBits(2, uint32_t(run)); break;
    case REPEAT_ZERO_3BITS: out.appendBits(3, uint32_t(run)); break;
    case REPEAT_ZERO_7BITS: out.appendBits(7, uint32_t(run)); break;
    default: break;
  }
}

// CanonicalHuffman.buildCodeLengthTree (CanonicalHuffman.java:285-343)
void build_code_length_tree(BitOut& out, const int* textCodeLengths) {
  std::vector<int> tc(N_SYMBOLS_TOTAL), tr(N_SYMBOLS_TOTAL);
  int nT = length_encode(N_SYMBOLS_TOTAL, textCodeLengths, tc.data(), tr.data());
  int counts[SYMBOL_SET_SIZE + 1] = {0};
  counts[SYMBOL_SET_SIZE] = 1;  // end-of-text
  for (int i = 0; i < nT; i++) counts[tc[i]]++;
  int ctLengths[SYMBOL_SET_SIZE + 1];
  canon_tree_lengths(counts, SYMBOL_SET_SIZE + 1, ctLengths, nullptr);
  std::vector<Code> ctCodes;
  canonical_codes(ctLengths, SYMBOL_SET_SIZE + 1, ctCodes);
  int cc[SYMBOL_SET_SIZE + 1], cr[SYMBOL_SET_SIZE + 1];
  int nC = length_encode(SYMBOL_SET_SIZE + 1, ctLengths, cc, cr);
  out.appendBit(0);  // reserved
  for (int i = 0; i < nC; i++) {  // LengthEncoder.writeEncodedLengths :169-195
    out.appendBits(5, uint32_t(cc[i]));
    write_run_extra(out, cc[i], cr[i]);
  }
  for (int i = 0; i < nT; i++) {
    write_code(out, ctCodes[tc[i]]);
    if (tc[i] > MAX_STANDARD_SYMBOL) write_run_extra(out, tc[i], tr[i]);
  }
}

// CanonicalHuffman.countSymbols (CanonicalHuffman.java:352-418)
void count_symbols(int n, const int32_t* text, int* counts) {
  for (int i = 0; i < N_SYMBOLS_TOTAL; i++) counts[i] = 0;
  counts[I_END_OF_TEXT] = 1;
  for (int i = 0; i < n; i++) {
    int32_t s = text[i];
    if (-128 <= s && s <= 127) counts[s + 128]++;
    else if (-512 <= s && s <= 511) { counts[I_ESCAPE_2BITS]++; counts[(s >> 2) + 128]++; }
    else if (-2048 <= s && s <= 2047) { counts[I_ESCAPE_2BITS] += 2; counts[(s >> 4) + 128]++; }
    else if (-8192 <= s && s <= 8191) { counts[I_ESCAPE_2BITS] += 3; counts[(s >> 6) + 128]++; }
    else if (-32768 <= s && s <= 32767) { counts[I_ESCAPE_1BYTE]++; counts[(s >> 8) + 128]++; }
    else if (s == INT4_NULL_CODE) counts[I_NULL_DATA_CODE]++;
    else if (-8388608 <= s && s <= 8388607) { counts[I_ESCAPE_1BYTE] += 2; counts[(s >> 16) + 128]++; }
    else { counts[I_ESCAPE_1BYTE] += 3; counts[(s >> 24) + 128]++; }
  }
}
}  // namespace

// CanonicalHuffman.encode(BitOutputStore, n, offset=0, text) (CanonicalHuffman.java:177-283)
void canon_encode(BitOut& out, int nSymbols, const int32_t* text) {
  if (nSymbols <= 0 || !text) throw std::invalid_argument("Empty or null data input data");
  int counts[N_SYMBOLS_TOTAL];
  count_symbols(nSymbols, text, counts);
  int lengths[N_SYMBOLS_TOTAL];
  canon_tree_lengths(counts, N_SYMBOLS_TOTAL, lengths, nullptr);
  std::vector<Code> codes;
  canonical_codes(lengths, N_SYMBOLS_TOTAL, codes);
  build_code_length_tree(out, lengths);
  auto sym = [&](int s) {
    if (codes[s].nBits == 0) throw std::runtime_error("canon_encode: symbol without a code (reference :258 vs :395 range bug)");
    write_code(out, codes[s]);
  };
  for (int i = 0; i < nSymbols; i++) {
    int32_t s = text[i];
    if (-128 <= s && s <= 127) { sym(s + 128); }
    else if (-512 <= s && s <= 511) {
      sym((s >> 2) + 128); sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t(s & 3));
    } else if (-2048 <= s && s <= 2047) {
      sym((s >> 4) + 128);
      sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t((s >> 2) & 3));
      sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t(s & 3));
    } else if (-8192 <= s && s <= 8191) {
      sym((s >> 6) + 128);
      sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t((s >> 4) & 3));
      sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t((s >> 2) & 3));
      sym(I_ESCAPE_2BITS); out.appendBits(2, uint32_t(s & 3));
    } else if (-32768 <= s && s <= 32767) {
      sym((s >> 8) + 128); sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t(s & 0xff));
    } else if (s == INT4_NULL_CODE) {
      sym(I_NULL_DATA_CODE);
    } else if (-8333608 <= s && s <= 8388607) {  // sic: -8333608, CanonicalHuffman.java:258
      sym((s >> 16) + 128);
      sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t((s >> 8) & 0xff));
      sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t(s & 0xff));
    } else {
      sym((s >> 24) + 128);
      sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t((s >> 16) & 0xff));
      sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t((s >> 8) & 0xff));
      sym(I_ESCAPE_1BYTE); out.appendBits(8, uint32_t(s & 0xff));
    }
  }
  sym(I_END_OF_TEXT);
}

namespace {
// CanonHuffTreeDecoder (CanonHuffTreeDecoder.java:68-129): explicit binary tree from code lengths.
struct TreeDecoder {
  std::vector<int> nodeIndex;
  int kLookup = 0;
  std::vector<int> lookup;
  explicit TreeDecoder(const std::vector<int>& symbolLengths) {
    int nSymbols = int(symbolLengths.size());
    std::vector<Code> codes;
    canonical_codes(symbolLengths.data(), nSymbols, codes);
    std::vector<int> order;
    for (int i = 0; i < nSymbols; i++) if (symbolLengths[i] > 0) order.push_back(i);
    if (order.empty()) throw std::runtime_error("canonical tree without symbols");  // Java: ArrayIndexOutOfBounds
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      if (symbolLengths[a] != symbolLengths[b]) return symbolLengths[a] < symbolLengths[b];
      return a < b;
    });
    int n = nSymbols * 2 + 2;
    nodeIndex.assign(size_t(n) * 3, -1);
    int nUsed = 3;
    int minLen = symbolLengths[order[0]];
    kLookup = minLen > 8 ? 8 : minLen;
    lookup.assign(size_t(1) << kLookup, 0);
    for (int s : order) {
      int index = 0, iLookup = 0;
      int len = codes[s].nBits;
      for (int k = 0; k < len; k++) {
        int bit = int((codes[s].bits >> (len - 1 - k)) & 1);
        iLookup |= (bit << k);
        int test = nodeIndex.at(index + 1 + bit);
        if (test < 0) { nodeIndex.at(index + 1 + bit) = nUsed; index = nUsed; nUsed += 3; }
        else index = test;
        if (k == kLookup - 1) lookup[iLookup] = index;
      }
      nodeIndex.at(index) = s;
    }
  }
  int walk(BitIn& in, int offset) const {
    while (nodeIndex.at(offset) == -1) {
      int nxt = nodeIndex.at(offset + 1 + in.getBit());
      if (nxt < 0) throw std::runtime_error("invalid canonical code");  // Java: ArrayIndexOutOfBounds(-1)
      offset = nxt;
    }
    return nodeIndex[offset];
  }
};
}  // namespace

// CanonicalHuffman.decode (:441-466) + decodeText (:469-519) + CanonHuffTreeDecoder.decodeTree (:131-177)
// + LengthEncoder.readEncodedLengths (:197-236)
bool canon_decode(BitIn& in, int nSymbolsInText, int32_t* text) {
  if (nSymbolsInText <= 0) return false;
  in.getBit();  // reserved
  std::vector<int> ctLengths(SYMBOL_SET_SIZE + 1, 0);
  {
    int k = 0, prior = 0;
    const int nS = SYMBOL_SET_SIZE + 1;
    auto put = [&](int v) { if (k >= nS) throw std::runtime_error("code-table lengths overrun"); ctLengths[k++] = v; };
    while (k < nS) {
      int index = int(in.getBits(5));
      if (index <= MAX_STANDARD_SYMBOL) { prior = index; put(index); }
      else if (index == REPEAT_PREV_2BITS) { int n = int(in.getBits(2)) + 3; for (int i = 0; i < n; i++) put(prior); }
      else if (index == REPEAT_ZERO_3BITS) { prior = 0; int n = int(in.getBits(3)) + 3; for (int i = 0; i < n; i++) put(0); }
      else if (index == REPEAT_ZERO_7BITS) { prior = 0; int n = int(in.getBits(7)) + 11; for (int i = 0; i < n; i++) put(0); }
      // other 5-bit values: the reference ignores them (infinite-loop guard is the bit store running dry)
    }
  }
  TreeDecoder codeTable(ctLengths);
  std::vector<int> textLengths(N_SYMBOLS_TOTAL + 1, 0);  // CanonicalHuffman.java:456 (one spare element)
  {
    int prior = 0;
    for (int i = 0; i < N_SYMBOLS_TOTAL; i++) {
      int start = codeTable.nodeIndex.at(1 + in.getBit());
      if (start < 0) throw std::runtime_error("invalid canonical code");
      int test = codeTable.walk(in, start);
      if (test <= MAX_STANDARD_SYMBOL) { textLengths[i] = test; prior = test; }
      else {
        int n = 0, val = 0;
        if (test == REPEAT_PREV_2BITS) { n = int(in.getBits(2)) + 3; val = prior; }
        else if (test == REPEAT_ZERO_3BITS) { prior = 0; n = int(in.getBits(3)) + 3; }
        else if (test == REPEAT_ZERO_7BITS) { prior = 0; n = int(in.getBits(7)) + 11; }
        else continue;  // EOT symbol of the code table: reference `default: break` leaves symbols[i] = 0
        for (int j = 0; j < n; j++) {
          if (i + j > N_SYMBOLS_TOTAL) throw std::runtime_error("text lengths overrun");
          textLengths[i + j] = val;
        }
        i += n - 1;
      }
    }
  }
  TreeDecoder textTree(textLengths);
  int32_t prior = 0;
  int iSymbol = 0;
  while (true) {  // decodeText :469-519 -- runs until end-of-text
    uint32_t iX = in.getBits(textTree.kLookup);
    int offset = textTree.lookup[iX];
    int symbol = textTree.walk(in, offset);
    if (symbol == I_END_OF_TEXT) break;
    if (symbol < 256) {
      symbol -= 128;
      if (iSymbol >= nSymbolsInText) throw std::out_of_range("canonical text overrun");
      text[iSymbol++] = symbol;
      prior = symbol;
    } else if (symbol == I_ESCAPE_2BITS) {
      uint32_t part = in.getBits(2);
      prior = int32_t((uint32_t(prior) << 2) | part);
      if (iSymbol < 1) throw std::out_of_range("escape before first symbol");
      text[iSymbol - 1] = prior;
    } else if (symbol == I_ESCAPE_1BYTE) {
      uint32_t part = in.getBits(8);
      prior = int32_t((uint32_t(prior) << 8) | part);
      if (iSymbol < 1) throw std::out_of_range("escape before first symbol");
      text[iSymbol - 1] = prior;
    } else if (symbol == I_NULL_DATA_CODE) {
      prior = INT4_NULL_CODE;
      if (iSymbol >= nSymbolsInText) throw std::out_of_range("canonical text overrun");
      text[iSymbol++] = INT4_NULL_CODE;
    }
  }
  return true;
}

}  // namespace g4o
