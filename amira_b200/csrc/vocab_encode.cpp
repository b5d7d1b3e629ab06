// vocab_encode.cpp -- host-side gene-call token parsing for libamira_gmg.so.
// Replaces Gene.__init__ / split_gene_and_strand (amira/construct_gene.py:48-67) and convert_genes
// (amira/construct_read.py:5-8): "+name" / "-name" -> strand * rank, with the same three failure
// modes (blank token, bad strand character, empty name).
#include <stdint.h>

#include <string>
#include <string_view>
#include <unordered_map>

#include "../../include/amira_gmg.h"

namespace amira {
void set_error(const char *fmt, ...);
}

extern "C" int amira_vocab_encode(const char *tokens_utf8, const int64_t *tok_off, int64_t n_tok,
                                  const char *vocab_utf8, const int64_t *vocab_off, int32_t n_vocab,
                                  int32_t *out_signed_ids, int64_t *bad_token) {
    if (n_tok < 0 || n_vocab < 0 || (n_tok > 0 && (!tokens_utf8 || !out_signed_ids)) ||
        (n_vocab > 0 && (!vocab_utf8 || !vocab_off))) {
        amira::set_error("bad arguments to amira_vocab_encode");
        return AMIRA_E_ARG;
    }
    std::unordered_map<std::string_view, int32_t> rank;
    rank.reserve((size_t)n_vocab * 2 + 16);
    for (int32_t v = 0; v < n_vocab; ++v)
        rank.emplace(std::string_view(vocab_utf8 + vocab_off[v], (size_t)(vocab_off[v + 1] - vocab_off[v])), v + 1);
    std::string name;
    // tok_off == NULL: the tokens are separated by '\n' and the blob is NUL-terminated (saves the caller a
    // per-token length pass); exactly n_tok tokens must be there
    const char *cursor = tokens_utf8;
    for (int64_t t = 0; t < n_tok; ++t) {
        const char *s;
        size_t len;
        if (tok_off) {
            s = tokens_utf8 + tok_off[t];
            len = (size_t)(tok_off[t + 1] - tok_off[t]);
        } else {
            s = cursor;
            const char *e = s;
            while (*e && *e != '\n') ++e;
            len = (size_t)(e - s);
            if (!*e && t + 1 < n_tok) {
                amira::set_error("token blob holds fewer than %lld tokens", (long long)n_tok);
                return AMIRA_E_ARG;
            }
            cursor = *e ? e + 1 : e;
        }
        auto fail = [&](int code, const char *what) {
            if (bad_token) *bad_token = t;
            if (code == AMIRA_E_BLANK_GENE) amira::set_error("%s", what);
            else amira::set_error("%s%.*s", what, (int)len, s);
            return code;
        };
        bool blank = true;
        for (size_t i = 0; i < len; ++i)
            if (s[i] != ' ') {
                blank = false;
                break;
            }
        if (blank) return fail(AMIRA_E_BLANK_GENE, "Gene information is missing");
        if (s[0] != '+' && s[0] != '-') return fail(AMIRA_E_BAD_STRAND, "Strand information missing for: ");
        if (len == 1) return fail(AMIRA_E_EMPTY_NAME, "Gene name information missing for: ");
        name.assign(s + 1, len - 1);
        for (char &c : name)
            if (c == ' ') c = '_';
        auto it = rank.find(std::string_view(name));
        if (it == rank.end()) return fail(AMIRA_E_UNKNOWN_GENE, "gene not in vocabulary: ");
        out_signed_ids[t] = s[0] == '+' ? it->second : -it->second;
    }
    if (!tok_off && n_tok > 0 && *cursor) {
        amira::set_error("token blob holds more than %lld tokens", (long long)n_tok);
        return AMIRA_E_ARG;
    }
    return AMIRA_OK;
}
