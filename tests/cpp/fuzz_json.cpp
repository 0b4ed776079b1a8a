// TEST INFRASTRUCTURE.  Drives crane_b200/csrc/json_min.h (the parser behind config.json and safetensors headers, i.e. untrusted
// input) with mutated documents under AddressSanitizer-free plain g++: every input must either parse or throw std::runtime_error --
// no crash, no hang, no read past the terminating NUL (the inputs are placed at the very end of an exactly-sized heap block so an
// over-read of more than a few bytes faults under glibc's allocator checks / valgrind; the depth and escape paths are what the
// advisor's round-1 findings named).
#include "../../crane_b200/csrc/json_min.h"

#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

static int run_one(const std::string& doc) {
    std::vector<char> buf(doc.size() + 1);          // exactly sized: the NUL is the last byte of the allocation
    memcpy(buf.data(), doc.data(), doc.size());
    buf[doc.size()] = 0;
    try {
        cbjson::Parser p(buf.data());
        cbjson::Value v = p.parse();
        (void)v;
        return 0;
    } catch (const std::runtime_error&) {
        return 1;
    }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    const std::string seeds[] = {
        R"({"model_type":"qwen3","hidden_size":1024,"rope_parameters":{"rope_theta":1e6,"mrope_section":[11,11,10]},"tie_word_embeddings":true,"x":null})",
        R"({"__metadata__":{"format":"pt"},"model.embed_tokens.weight":{"dtype":"BF16","shape":[151936,1024],"data_offsets":[0,311164928]}})",
        R"(["a\u00e9\ud83d\ude00b","\\\"\/\b\f\n\r\t",-1.5e-3,0,true,false,[[[[]]]],{}])",
    };
    // hand-written hostile inputs first
    const std::string hostile[] = {
        "", "{", "[", "\"", "\"\\", "\"\\u", "\"\\u12", "\"\\ud83d", "\"\\ud83d\\u", "{\"a\":", "{\"a\"", "{\"a\":1,", "[1,", "-", "1e", "tru", "nul",
        std::string(100000, '['), std::string(100000, '{'), "[" + std::string(5000, ' ') + "]", "\"" + std::string(70000, 'x'),
        "{\"a\":1}garbage", "\x01\x02\x03", "{\"k\":\"\\u0000\"}", "[1e999999,-1e999999]",
    };
    int parsed = 0, rejected = 0;
    for (const auto& h : hostile) (run_one(h) ? rejected : parsed)++;
    std::mt19937 rng(12345);
    for (int it = 0; it < iters; ++it) {
        std::string d = seeds[it % 3];
        const int nmut = 1 + rng() % 4;
        for (int m = 0; m < nmut && !d.empty(); ++m) {
            const size_t pos = rng() % d.size();
            switch (rng() % 5) {
                case 0: d[pos] = (char)(rng() % 256); break;                       // byte flip (NUL included: truncates the document)
                case 1: d.erase(pos, 1 + rng() % 8); break;                        // delete a run
                case 2: d.insert(pos, 1, "\"\\{}[],:u0"[rng() % 10]); break;       // insert a structural character
                case 3: d.resize(pos); break;                                      // truncate
                case 4: d.insert(pos, d.substr(pos, rng() % 16)); break;           // duplicate a run
            }
        }
        (run_one(d) ? rejected : parsed)++;
    }
    printf("json fuzz ok: %d parsed, %d rejected\n", parsed, rejected);
    return 0;
}
