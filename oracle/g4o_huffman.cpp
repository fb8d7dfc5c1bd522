// CPU ORACLE (test infrastructure only) -- legacy byte-alphabet Huffman coder.
// Restates C/compress/HuffmanEncoder.java:124-305 and C/compress/HuffmanDecoder.java:65-187
// (C/ = /root/reference/core/src/main/java/org/gridfour/).
#include "g4oracle.h"
#include <algorithm>

namespace g4o {

namespace {
struct Node {
  bool isLeaf = false;
  int symbol = -1;
  int count = 0;
  int bit = 0;
  int next = -1, left = -1, right = -1;
  int nBitsInCode = 0;
  std::vector<uint8_t> path;  // root->leaf bits in path order (leaves only)
};

// Builds the reference's tree.  nodes[0..255] are the leaves (index == symbol).  Returns the root
// index, or -1 for the single-symbol case (then *single = that symbol).  HuffmanEncoder.java:128-194
int build_tree(std::vector<Node>& nodes, int nSymbols, const uint8_t* symbols, int* nLeaf, int* single) {
  nodes.assign(256, Node());
  for (int i = 0; i < 256; i++) { nodes[i].isLeaf = true; nodes[i].symbol = i; }
  for (int i = 0; i < nSymbols; i++) nodes[symbols[i]].count++;
  int sortIdx[256];
  for (int i = 0; i < 256; i++) sortIdx[i] = i;
  // compareTo: count ascending, then symbol ascending (HuffmanEncoder.java:86-92)
  std::sort(sortIdx, sortIdx + 256, [&](int a, int b) {
    if (nodes[a].count != nodes[b].count) return nodes[a].count < nodes[b].count;
    return nodes[a].symbol < nodes[b].symbol;
  });
  int firstIndex = -1;
  for (int i = 0; i < 256; i++) if (nodes[sortIdx[i]].count > 0) { firstIndex = i; break; }
  if (firstIndex == 255) { *single = nodes[sortIdx[255]].symbol; *nLeaf = 1; return -1; }
  if (firstIndex < 0) throw std::invalid_argument("huffman_encode: no symbols");  // Java: ArrayIndexOutOfBounds
  for (int i = firstIndex; i < 255; i++) nodes[sortIdx[i]].next = sortIdx[i + 1];
  *nLeaf = 256 - firstIndex;
  int first = sortIdx[firstIndex];
  int root = -1;
  while (true) {
    int left = first;
    int right = nodes[first].next;
    first = nodes[right].next;
    nodes[left].next = -1;
    nodes[right].next = -1;
    Node br;
    br.left = left; br.right = right;
    br.count = nodes[left].count + nodes[right].count;
    nodes[left].bit = 0; nodes[right].bit = 1;
    int b = int(nodes.size());
    nodes.push_back(br);
    if (first < 0) { root = b; break; }
    if (nodes[first].count >= nodes[b].count) {
      nodes[b].next = first;
      first = b;
    } else {
      int node = nodes[first].next;
      int prior = first;
      while (node >= 0 && nodes[node].count < nodes[b].count) { prior = node; node = nodes[node].next; }
      nodes[prior].next = b;
      if (node >= 0) nodes[b].next = node;
    }
  }
  return root;
}

// Pre-order traversal: writes the tree when `out` is given and records each leaf's path.
// HuffmanEncoder.java:221-305
void walk_tree(std::vector<Node>& nodes, int root, int nLeaf, BitOut* out) {
  if (out) out->appendBits(8, uint32_t(nLeaf - 1));
  std::vector<int> path(257), branch(257);
  path[0] = root; branch[0] = 0;
  int depth = 1;
  while (depth > 0) {
    int index = depth - 1;
    int p = path[index];
    switch (branch[index]) {
      case 0:
        if (nodes[p].isLeaf) {
          if (out) { out->appendBit(1); out->appendBits(8, uint32_t(nodes[p].symbol)); }
          nodes[p].path.clear();
          for (int i = 1; i < depth; i++) nodes[p].path.push_back(uint8_t(nodes[path[i]].bit));
          nodes[p].nBitsInCode = depth - 1;
          depth--;
        } else {
          if (out) out->appendBit(0);
          branch[index] = 1;
          branch[depth] = 0;
          path[depth] = nodes[p].left;
          depth++;
        }
        break;
      case 1:
        branch[index] = 2;
        branch[depth] = 0;
        path[depth] = nodes[p].right;
        depth++;
        break;
      default:
        branch[index] = 0;
        depth--;
        break;
    }
  }
}
}  // namespace

void huffman_encode(BitOut& out, int nSymbols, const uint8_t* symbols) {
  std::vector<Node> nodes;
  int nLeaf = 0, single = -1;
  int root = build_tree(nodes, nSymbols, symbols, &nLeaf, &single);
  if (root < 0) {  // HuffmanEncoder.java:147-157
    out.appendBits(8, 0);
    out.appendBit(1);
    out.appendBits(8, uint32_t(single));
    return;
  }
  walk_tree(nodes, root, nLeaf, &out);
  for (int i = 0; i < nSymbols; i++) {  // HuffmanEncoder.java:198-213
    const Node& n = nodes[symbols[i]];
    for (uint8_t b : n.path) out.appendBit(b);
  }
}

int huffman_code_lengths(int nSymbols, const uint8_t* symbols, int lengths[256]) {
  std::vector<Node> nodes;
  int nLeaf = 0, single = -1;
  for (int i = 0; i < 256; i++) lengths[i] = 0;
  int root = build_tree(nodes, nSymbols, symbols, &nLeaf, &single);
  if (root < 0) return 1;
  walk_tree(nodes, root, nLeaf, nullptr);
  for (int i = 0; i < 256; i++) lengths[i] = nodes[i].count > 0 ? nodes[i].nBitsInCode : 0;
  return nLeaf;
}

// HuffmanDecoder.java:65-187
void huffman_decode(BitIn& in, int nSymbols, uint8_t* symbols) {
  int nLeafsToDecode = int(in.getBits(8)) + 1;
  int rootBit = in.getBit();
  if (rootBit == 1) {
    uint8_t s = uint8_t(in.getBits(8));
    for (int i = 0; i < nSymbols; i++) symbols[i] = s;
    return;
  }
  std::vector<int> nodeIndex(size_t(nLeafsToDecode) * 6, 0);
  std::vector<int> stack(size_t(nLeafsToDecode) + 1, 0);
  int iStack = 0;
  nodeIndex[0] = -1;
  int nodeIndexCount = 3;
  int nLeafsDecoded = 0;
  while (nLeafsDecoded < nLeafsToDecode) {
    int offset = stack.at(iStack);
    if (nodeIndex.at(offset + 1) == 0) nodeIndex.at(offset + 1) = nodeIndexCount;
    else nodeIndex.at(offset + 2) = nodeIndexCount;
    int bit = in.getBit();
    if (bit == 1) {
      nLeafsDecoded++;
      nodeIndex.at(nodeIndexCount++) = int(in.getBits(8));
      nodeIndex.at(nodeIndexCount++) = 0;
      nodeIndex.at(nodeIndexCount++) = 0;
      if (nLeafsDecoded == nLeafsToDecode) break;
      while (nodeIndex.at(offset + 2) != 0) {
        iStack--;
        offset = stack.at(iStack);  // throws on a malformed tree (Java: ArrayIndexOutOfBounds)
      }
    } else {
      iStack++;
      stack.at(iStack) = nodeIndexCount;
      nodeIndex.at(nodeIndexCount++) = -1;
      nodeIndex.at(nodeIndexCount++) = 0;
      nodeIndex.at(nodeIndexCount++) = 0;
    }
  }
  for (int i = 0; i < nSymbols; i++) {
    int offset = nodeIndex.at(1 + in.getBit());
    while (nodeIndex.at(offset) == -1) offset = nodeIndex.at(offset + 1 + in.getBit());
    symbols[i] = uint8_t(nodeIndex[offset]);
  }
}

}  // namespace g4o
