// C++ port of the reference's own tests (huff_coding/tests/{comp_decomp,tree_init}.rs and the comp.rs doctests)
// against include/huff_coding.hpp.  Modes: "host" (tree / container only, no GPU) and "gpu" (everything).
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/huff_coding.hpp"

using namespace huff_coding;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<uint8_t> bytes(const char *s) { return std::vector<uint8_t>(s, s + std::strlen(s)); }

static int host_tests() {
    // tests/tree_init.rs::tree_normal_init
    auto t = HuffTree::from_weights({{0, 5}, {1, 9}, {2, 12}, {3, 13}, {4, 16}, {5, 45}});
    auto c = t.read_codes();
    CHECK(c[0] == "1100" && c[1] == "1101" && c[2] == "100" && c[3] == "101" && c[4] == "111" && c[5] == "0");
    // tests/tree_init.rs::tree_single_branch
    auto s = HuffTree::from_weights({{0xF4, 78}}).read_codes();
    CHECK(s.size() == 1 && s[0xF4] == "0");
    // tests/tree_init.rs::tree_invalid_weights
    bool panicked = false;
    try { HuffTree::from_weights({}); } catch (const Panic &e) { panicked = std::string(e.what()) == "provided empty weights"; }
    CHECK(panicked);
    // comp.rs:219-262: container blob of compress(b"abbccc")
    CompressData cd({0xBC, 0x00}, 7, HuffTree::from_weights({{'a', 1}, {'b', 2}, {'c', 3}}));
    auto blob = cd.to_bytes();
    const uint8_t want[] = {0x37, 0, 0, 0, 4, 0x98, 0xE6, 0x13, 0x10, 0xBC, 0x00};
    CHECK(blob.size() == sizeof want && std::memcmp(blob.data(), want, sizeof want) == 0);
    auto back = CompressData::try_from_bytes(blob);
    CHECK(back.padding_bits() == 7 && back.comp_bytes() == cd.comp_bytes());
    return 0;
}

static int gpu_tests() {
    // tests/comp_decomp.rs::compress_decompress
    auto text = bytes("float Q_rsqrt( float number )\n{\n\tlong i;\n\tfloat x2, y;\n\tconst float threehalfs = 1.5F;\n"
                      "\tx2 = number * 0.5F;\n\ty  = number;\n\ti  = * ( long * ) &y; // evil floating point bit level hacking\n"
                      "\ti  = 0x5f3759df - ( i >> 1 ); // what the fuck?\n\ty  = * ( float * ) &i;\n\treturn y;\n}");
    auto cd = compress(text);
    CHECK(decompress(cd) == text);
    // comp.rs doctest: compress(b"abbccc").to_bytes()
    auto blob = compress(bytes("abbccc")).to_bytes();
    const uint8_t want[] = {0x37, 0, 0, 0, 4, 0x98, 0xE6, 0x13, 0x10, 0xBC, 0x00};
    CHECK(blob.size() == sizeof want && std::memcmp(blob.data(), want, sizeof want) == 0);
    CHECK(decompress(CompressData::try_from_bytes(blob)) == bytes("abbccc"));
    // comp.rs:399-415: missing letter
    bool err = false;
    try { compress_with_tree(bytes("abbccc"), HuffTree::from_weights(build_weights_map(bytes("abb")))); }
    catch (const CompressError &e) { err = e.missing_letter() == 'c'; }
    CHECK(err);
    // weights.rs:64-68
    auto w = build_weights_map(bytes("aabbbc"));
    CHECK(w['a'] == 2 && w['b'] == 3 && w['c'] == 1);
    // empty input
    bool panicked = false;
    try { compress({}); } catch (const Panic &) { panicked = true; }
    CHECK(panicked);
    return 0;
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    int rc = host_tests();
    if (rc == 0 && gpu) rc = gpu_tests();
    if (rc == 0) std::printf("cpp tests ok (%s)\n", gpu ? "host+gpu" : "host");
    return rc;
}
