// huff_coding.hpp -- C++ host-side mirror of `huff_coding::prelude` for the u8 alphabet over the C ABI of
// libhuffb200.so.  Same names, argument meaning and error behaviour as the reference (the reference is compiled Rust;
// no Rust toolchain exists in this image, so the compiled-language host side is C++).  Header-only.
//
//   build_weights_map            huff_coding/src/weights.rs:82-84
//   HuffTree::from_weights       huff_coding/src/tree/tree_inner.rs:281-320
//   HuffTree::read_codes         huff_coding/src/tree/tree_inner.rs:356-419
//   compress / compress_with_tree / decompress   huff_coding/src/comp.rs:353-356 / 419-451 / 487-519
//   CompressData                 huff_coding/src/comp.rs:41-89, to_bytes :279-300, try_from_bytes :128-184
//
// Reference panics are thrown as huff_coding::Panic with the reference's message.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "huffb200.h"

namespace huff_coding {

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

// comp.rs:561-590
struct CompressError : std::runtime_error {
    uint8_t missing;
    explicit CompressError(uint8_t m) : std::runtime_error("letter not found in codes"), missing(m) {}
    const char *message() const { return what(); }
    uint8_t missing_letter() const { return missing; }
};
// comp.rs:531-554
struct CompressedDataFromBytesError : std::runtime_error { using std::runtime_error::runtime_error; };

namespace detail {
inline hb_ctx *ctx() {
    thread_local std::unique_ptr<hb_ctx, hb_status (*)(hb_ctx *)> holder(nullptr, hb_ctx_destroy);
    if (!holder) {
        hb_ctx *c = nullptr;
        if (hb_ctx_create(0, &c) != HB_OK) throw CudaError(std::string("libhuffb200: ") + hb_last_error() + " (no CPU fallback)");
        holder.reset(c);
    }
    return holder.get();
}
inline void check(hb_status s, uint8_t missing = 0) {
    switch (s) {
        case HB_OK: return;
        case HB_ERR_EMPTY_WEIGHTS: throw Panic("provided empty weights");
        case HB_ERR_EMPTY_COMP: throw Panic("provided comp_bytes are empty");
        case HB_ERR_BAD_PADDING: throw Panic("padding bits cannot be larger than 7");
        case HB_ERR_TREE_LEN: throw Panic("stored tree length must be at least 2");
        case HB_ERR_MISSING_LETTER: throw CompressError(missing);
        case HB_ERR_INVALID_TREE: throw CompressedDataFromBytesError("invalid tree in slice");
        case HB_ERR_BYTES_SHORT: throw CompressedDataFromBytesError("slice too short");
        case HB_ERR_CUDA: throw CudaError(hb_last_error());
        default: throw std::runtime_error(hb_status_str(s));
    }
}
}  // namespace detail

// weights.rs:82-84 (ordered map: ascending letters = the canonical leaf order)
inline std::map<uint8_t, size_t> build_weights_map(const std::vector<uint8_t> &letters) {
    uint64_t w[256];
    detail::check(hb_histogram_u8(detail::ctx(), letters.data(), letters.size(), w));
    std::map<uint8_t, size_t> m;
    for (int b = 0; b < 256; b++) if (w[b]) m[static_cast<uint8_t>(b)] = static_cast<size_t>(w[b]);
    return m;
}

// tree_inner.rs:193-196
class HuffTree {
public:
    static HuffTree from_weights(const std::map<uint8_t, size_t> &weights) {
        std::vector<uint8_t> letters;
        std::vector<uint64_t> ws;
        for (auto &kv : weights) { letters.push_back(kv.first); ws.push_back(kv.second); }
        HuffTree t;
        detail::check(hb_tree_from_pairs(letters.data(), ws.data(), letters.size(), t.raw_.get()));
        return t;
    }
    // letter -> code as a '0'/'1' string (bitvec Msb0 order)
    std::map<uint8_t, std::string> read_codes() const {
        std::map<uint8_t, std::string> out;
        for (int b = 0; b < 256; b++) {
            if (!raw_->has_code[b]) continue;
            std::string s;
            for (int k = raw_->code_len[b] - 1; k >= 0; k--) s.push_back(((raw_->code[b] >> k) & 1) ? '1' : '0');
            out[static_cast<uint8_t>(b)] = s;
        }
        return out;
    }
    const hb_tree *raw() const { return raw_.get(); }
    hb_tree *raw() { return raw_.get(); }
    HuffTree() : raw_(new hb_tree()) {}
    HuffTree(const HuffTree &o) : raw_(new hb_tree(*o.raw_)) {}
    HuffTree(HuffTree &&) = default;
    HuffTree &operator=(HuffTree &&) = default;
private:
    std::unique_ptr<hb_tree> raw_;
};

// comp.rs:41-89
class CompressData {
public:
    CompressData(std::vector<uint8_t> comp_bytes, uint8_t padding_bits, HuffTree tree)
        : comp_(std::move(comp_bytes)), pad_(padding_bits), tree_(std::move(tree)) {
        if (comp_.empty()) throw Panic("provided comp_bytes are empty");
        if (pad_ > 7) throw Panic("padding bits cannot be larger than 7");
    }
    const std::vector<uint8_t> &comp_bytes() const { return comp_; }
    uint8_t padding_bits() const { return pad_; }
    const HuffTree &huff_tree() const { return tree_; }
    std::vector<uint8_t> to_bytes() const {                                  // comp.rs:279-300
        std::vector<uint8_t> out(comp_.size() + 512);
        size_t n = 0;
        detail::check(hb_to_bytes(comp_.data(), comp_.size(), pad_, tree_.raw(), out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }
    static CompressData try_from_bytes(const std::vector<uint8_t> &bytes) {   // comp.rs:128-184
        HuffTree t;
        size_t off = 0, len = 0;
        uint8_t pad = 0;
        detail::check(hb_try_from_bytes(bytes.data(), bytes.size(), t.raw(), &off, &len, &pad));
        return CompressData(std::vector<uint8_t>(bytes.begin() + off, bytes.begin() + off + len), pad, std::move(t));
    }
private:
    std::vector<uint8_t> comp_;
    uint8_t pad_;
    HuffTree tree_;
};

// comp.rs:353-356
inline CompressData compress(const std::vector<uint8_t> &letters) {
    HuffTree t;
    uint8_t *p = nullptr, pad = 0;
    size_t n = 0;
    detail::check(hb_compress_u8(detail::ctx(), letters.data(), letters.size(), HB_ORDER_ASC, t.raw(), &p, &n, &pad));
    std::vector<uint8_t> v(p, p + n);
    hb_free(p);
    return CompressData(std::move(v), pad, std::move(t));
}

// comp.rs:419-451
inline CompressData compress_with_tree(const std::vector<uint8_t> &letters, HuffTree tree) {
    uint8_t *p = nullptr, pad = 0, missing = 0;
    size_t n = 0;
    const hb_status st = hb_compress_with_tree_u8(detail::ctx(), letters.data(), letters.size(), tree.raw(), &p, &n, &pad, &missing);
    detail::check(st, missing);                              // (two statements: `missing` must be read after the call)
    std::vector<uint8_t> v(p, p + n);
    hb_free(p);
    return CompressData(std::move(v), pad, std::move(tree));
}

// comp.rs:487-519
inline std::vector<uint8_t> decompress(const CompressData &cd) {
    uint8_t *p = nullptr;
    size_t n = 0;
    detail::check(hb_decompress_u8(detail::ctx(), cd.comp_bytes().data(), cd.comp_bytes().size(), cd.padding_bits(),
                                   cd.huff_tree().raw(), &p, &n));
    std::vector<uint8_t> v(p, p + n);
    hb_free(p);
    return v;
}

}  // namespace huff_coding
