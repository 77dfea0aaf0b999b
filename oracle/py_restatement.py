"""Second, structurally independent restatement of the reference path in pure Python -- TEST INFRASTRUCTURE.

It follows the Rust control flow literally (boxed branches, a list-backed BinaryHeap with std's hole moves, the
ByteWeights iterator transcribed statement by statement) so that a transcription slip in oracle/huff_oracle.c and a
slip here are unlikely to coincide.  Small inputs only (pure-Python loops).

Reference anchors (paths relative to /root/reference/huff_coding/src):
  weights.rs:116-123, 265-279, 396-415   tree/branch_heap.rs:24-83   tree/leaf.rs:31-47
  tree/tree_inner.rs:281-320, 388-440, 632-668   comp.rs:419-451, 487-519
"""
from __future__ import annotations


# ---------------------------------------------------------------- weights
def build_weights_map(letters) -> dict:
    """weights.rs:116-123.  A dict keeps first-insertion order; Rust's HashMap order is random per process."""
    m: dict = {}
    for l in letters:
        m[l] = m.get(l, 0) + 1
    return m


class ByteWeights:
    """weights.rs:175-443"""

    def __init__(self, data: bytes = b""):
        self.weights = [0] * 256
        self.len = 0
        for byte in data:                      # :268-273
            if self.weights[byte] == 0:
                self.len += 1
            self.weights[byte] += 1

    def get(self, byte: int):                  # :323-329
        w = self.weights[byte]
        return None if w == 0 else w

    def is_empty(self):
        return self.len == 0

    def __iter__(self):                        # :396-415, statement by statement
        current_index = 0
        while True:
            if current_index == 256:
                return
            stop = False
            while self.get(current_index & 0xFF) is None:      # `current_index as u8`
                if current_index == 256:
                    stop = True
                    break
                current_index += 1
            if stop:
                return
            entry = (current_index & 0xFF, self.get(current_index & 0xFF))
            if current_index != 256:
                current_index += 1
            yield entry


# ---------------------------------------------------------------- std BinaryHeap (max-heap on `le`)
class BinaryHeap:
    """Rust std `alloc::collections::binary_heap` moves, generic over a `le(a, b)` predicate (a <= b)."""

    def __init__(self, le):
        self.data = []
        self.le = le

    def __len__(self):
        return len(self.data)

    def push(self, item):
        old_len = len(self.data)
        self.data.append(item)
        self._sift_up(0, old_len)

    def pop(self):
        item = self.data.pop()
        if self.data:
            item, self.data[0] = self.data[0], item
            self._sift_down_to_bottom(0)
        return item

    def _sift_up(self, start, pos):
        elem = self.data[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if self.le(elem, self.data[parent]):
                break
            self.data[pos] = self.data[parent]
            pos = parent
        self.data[pos] = elem
        return pos

    def _sift_down_to_bottom(self, pos):
        end = len(self.data)
        start = pos
        elem = self.data[pos]
        child = 2 * pos + 1
        while child <= max(end - 2, 0) and end >= 2:
            if self.le(self.data[child], self.data[child + 1]):
                child += 1
            self.data[pos] = self.data[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1:
            self.data[pos] = self.data[child]
            pos = child
        self.data[pos] = elem
        self._sift_up(start, pos)


# ---------------------------------------------------------------- tree
class Branch:
    """tree/branch.rs:158-162 + tree/leaf.rs:25-29"""
    __slots__ = ("letter", "weight", "code", "left", "right")

    def __init__(self, letter, weight, children=None):
        self.letter, self.weight, self.code = letter, weight, None
        self.left, self.right = children if children else (None, None)

    def has_children(self):
        return self.left is not None


def tree_from_weights(pairs) -> Branch:
    """tree_inner.rs:281-320.  `pairs` = iterable of (letter, weight) in heap insertion order."""
    pairs = list(pairs)
    if not pairs:
        raise ValueError("provided empty weights")
    # branch_heap.rs:67-71: item.cmp(other) = other.leaf.cmp(self.leaf)  =>  a <= b  <=>  b.weight <= a.weight
    heap = BinaryHeap(lambda a, b: b.weight <= a.weight)
    for l, f in pairs:
        heap.push(Branch(l, f))
    while len(heap) > 1:
        mn = heap.pop()
        nx = heap.pop()
        heap.push(Branch(None, mn.weight + nx.weight, (mn, nx)))
    root = heap.pop()
    if root.has_children():
        _set_codes(root, None)
    else:
        root.code = "0"
    return root


def _set_codes(parent: Branch, parent_code):
    """tree_inner.rs:422-440"""
    if parent.has_children():
        for pos, child in ((0, parent.left), (1, parent.right)):
            child.code = (parent_code or "") + ("1" if pos else "0")
            _set_codes(child, child.code)


def read_codes(root: Branch) -> dict:
    """tree_inner.rs:388-419"""
    codes: dict = {}

    def set_codes(node: Branch, pos_in_parent: bool):
        if node.has_children():
            for pos, child in enumerate((node.left, node.right)):
                if child.letter is not None:
                    codes[child.letter] = child.code
                else:
                    set_codes(child, pos != 0)
        else:
            codes[node.letter] = "1" if pos_in_parent else "0"

    if root.has_children():
        set_codes(root.left, False)
        set_codes(root.right, True)
    else:
        codes[root.letter] = "0"
    return codes


def as_bin(root: Branch) -> str:
    """tree_inner.rs:632-668 as a '0'/'1' string"""
    out = []

    def rec(node):
        if node.has_children():
            out.append("1")
            rec(node.left)
            rec(node.right)
        else:
            out.append("0" + format(node.letter, "08b"))

    import sys
    sys.setrecursionlimit(10000)
    rec(root)
    return "".join(out)


# ---------------------------------------------------------------- comp
def compress_with_tree(letters, root: Branch):
    """comp.rs:419-451 -> (bytes, padding_bits)"""
    codes = read_codes(root)
    out = bytearray()
    comp_byte, bit_ptr = 0, 7
    for letter in letters:
        if letter not in codes:
            raise KeyError(letter)
        for ch in codes[letter]:
            comp_byte |= (1 if ch == "1" else 0) << bit_ptr
            if bit_ptr == 0:
                out.append(comp_byte)
                comp_byte, bit_ptr = 0, 7
            else:
                bit_ptr -= 1
    padding_bits = 0 if bit_ptr == 7 else bit_ptr + 1
    if padding_bits != 0:
        out.append(comp_byte)
    if not out:
        raise ValueError("provided comp_bytes are empty")
    return bytes(out), padding_bits


def compress(data: bytes, order: str = "asc"):
    """comp.rs:353-356 with a fixed leaf order: 'asc' (canonical), 'first_seen' (dict order), 'byteweights'."""
    if order == "asc":
        pairs = sorted(build_weights_map(data).items())
    elif order == "first_seen":
        pairs = list(build_weights_map(data).items())
    elif order == "byteweights":
        pairs = list(ByteWeights(data))
    else:
        raise ValueError(order)
    root = tree_from_weights(pairs)
    comp, pad = compress_with_tree(data, root)
    return comp, pad, root


def decompress(comp: bytes, padding_bits: int, root: Branch) -> bytes:
    """comp.rs:487-519"""
    if not comp:
        raise ValueError("provided comp_bytes are empty")
    out = bytearray()
    cur = root
    for i, byte in enumerate(comp):
        nbits = 8 - padding_bits if i == len(comp) - 1 else 8
        for bit_ptr in range(nbits):
            if cur.has_children():
                cur = cur.right if (byte >> (7 - bit_ptr)) & 1 else cur.left
            if not cur.has_children():
                out.append(cur.letter)
                cur = root
    return bytes(out)


# ---------------------------------------------------------------- container format (independent of huff_oracle.c)
class FromBinError(Exception):
    """tree_inner.rs:673-700"""


class FromBytesError(Exception):
    """comp.rs:531-554 (CompressedDataFromBytesError)"""


class Panic(Exception):
    """a reference panic!()"""


def try_from_bin(bits: str) -> Branch:
    """tree_inner.rs:522-604, transcribed as the reference writes it: a recursive descent over an iterator of bits.
    `bits` is a '0'/'1' string (the BitVec)."""
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    it = iter(bits)

    def read_branch() -> Branch:
        bit = next(it, None)
        if bit is None:                                           # :531-536
            raise FromBinError("Provided BitVec is too small for an encoded HuffTree")
        if bit == "1":                                            # :538-546 joint branch: left first, then right
            left = read_branch()
            right = read_branch()
            return Branch(None, 0, (left, right))
        letter_bits = [b for _, b in zip(range(8), it)]           # :556 bits.take(size_of_bits::<u8>())
        if len(letter_bits) != 8:                                 # :557-561
            raise FromBinError("Provided BitVec is too small for an encoded HuffTree")
        byte, bit_ptr = 0, 7
        for b in letter_bits:                                     # :562-570
            byte |= int(b) << bit_ptr
            bit_ptr -= 1
        return Branch(byte, 0)

    root = read_branch()
    if next(it, None) is not None:                                # :587-591
        raise FromBinError("Provided BitVec is too big for an encoded HuffTree")
    if root.has_children():                                       # :595-600
        _set_codes(root, None)
    else:
        root.code = "0"
    return root


def to_bytes(comp: bytes, padding_bits: int, root: Branch) -> bytes:
    """comp.rs:279-300"""
    tree_bin = as_bin(root)
    tree_pad = (8 - len(tree_bin) % 8) % 8                        # utils.rs:37-40 calc_padding_bits
    tree_bytes_len = (len(tree_bin) + tree_pad) // 8
    padded = tree_bin + "0" * tree_pad                            # BitVec::into_vec: dead bits are zero
    out = bytearray([(tree_pad << 4) + padding_bits])
    out += tree_bytes_len.to_bytes(4, "big")
    out += bytes(int(padded[i:i + 8], 2) for i in range(0, len(padded), 8))
    out += bytes(comp)
    return bytes(out)


def try_from_bytes(blob: bytes):
    """comp.rs:128-184 -> (comp_bytes, padding_bits, root); FromBytesError for its Err values, Panic for its panics
    (tree length below 2, and CompressData::new's invariants, comp.rs:56-61)."""
    blob = bytes(blob)
    if len(blob) < 1:                                             # :143 bytes.get(0)
        raise FromBytesError("slice is empty")
    tree_padding_bits = blob[0] >> 4
    data_padding_bits = blob[0] & 0b0000_1111
    if len(blob) < 5:                                             # :148-152 bytes.get(1..5)
        raise FromBytesError("slice too short to read tree length")
    tree_len = int.from_bytes(blob[1:5], "big")
    if tree_len < 2:                                              # :153-155
        raise Panic("stored tree length must be at least 2")
    if 5 + tree_len > len(blob):                                  # :160 bytes.get(5..5 + tree_len)
        raise FromBytesError("slice too short to read tree")
    bits = "".join(format(b, "08b") for b in blob[5:5 + tree_len])
    for _ in range(tree_padding_bits):                            # :163 b.pop() (None on an empty vector)
        bits = bits[:-1]
    try:
        root = try_from_bin(bits)
    except FromBinError:
        raise FromBytesError("invalid tree in slice") from None   # :167-177
    comp = blob[5 + tree_len:]                                    # :180 (an empty tail is Some(&[]))
    if len(comp) == 0:                                            # comp.rs:56-58
        raise Panic("provided comp_bytes are empty")
    if data_padding_bits > 7:                                     # comp.rs:59-61
        raise Panic("padding bits cannot be larger than 7")
    return comp, data_padding_bits, root
