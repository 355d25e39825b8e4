// Host-side PDB text parser behind the C ABI (no device code): the gemmi-free replacement of the reference's
// read_pdb (src/structure_io.py:6-55, which wraps gemmi.read_pdb(path, max_line_length=80)).
//
// Semantics kept from the reference / gemmi:
//   * ATOM and HETATM records only, fixed columns of the PDB format, lines cut at 80 characters;
//   * one model per MODEL ... ENDMDL block (a file without MODEL records is model 0); reading stops at END;
//   * inside a model, the parts of a chain that are separated in the file (typically HETATM records after the
//     last TER) are moved behind the chain's first part, keeping file order inside the chain (gemmi's
//     merge_chain_parts) -- the reference iterates chain by chain (model.all());
//   * alternate locations: an atom WITH an altloc flag is dropped when an atom with the same
//     chain / residue number / atom name key (and an altloc flag) was seen before, in any model
//     (src/structure_io.py:24-31: the key has no model index);
//   * het flag 'A' / 'H' = record type of the atom's residue, element from columns 77-78 (title case), or, when
//     those are blank, from the atom name columns (right-justified two-letter names start in column 13).
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace {

struct Line {
    const char *p;
    int n;
};

inline bool starts(const Line &l, const char *tag) {
    const int t = (int)strlen(tag);
    return l.n >= t && memcmp(l.p, tag, t) == 0;
}

// columns [a, b) of a line, blank padded
inline void field(const Line &l, int a, int b, char *dst) {
    for (int i = a; i < b; ++i) dst[i - a] = i < l.n ? l.p[i] : ' ';
    dst[b - a] = 0;
}

inline void strip_copy(const char *src, char *dst, int width) {      // stripped, NUL padded to `width`
    int a = 0, b = (int)strlen(src);
    while (a < b && isspace((unsigned char)src[a])) ++a;
    while (b > a && isspace((unsigned char)src[b - 1])) --b;
    memset(dst, 0, width);
    memcpy(dst, src + a, std::min(b - a, width));
}

}  // namespace

extern "C" int pesto_pdb_count_atoms_host(const char *text, size_t len) {
    if (!text) return 0;
    int n = 0;
    for (size_t i = 0; i < len;) {
        const char *e = (const char *)memchr(text + i, '\n', len - i);
        const size_t ll = e ? (size_t)(e - (text + i)) : len - i;
        if (ll >= 6 && (memcmp(text + i, "ATOM  ", 6) == 0 || memcmp(text + i, "HETATM", 6) == 0)) ++n;
        i += ll + 1;
    }
    return n;
}

extern "C" int pesto_pdb_parse_host(const char *text, size_t len, int capacity, float *xyz, char *name4, char *element2,
                                    char *resname3, int32_t *resid, char *het, char *chain, int32_t *model, char *icode,
                                    float *bfactor, int *n_out) {
    using namespace pesto;
    if (!text || !xyz || !name4 || !element2 || !resname3 || !resid || !het || !chain || !model || !icode || !bfactor || !n_out) {
        set_error("pdb_parse: null pointer");
        return PESTO_EINVAL;
    }
    struct Atom {
        float x, y, z, b;
        char name[5], elem[3], resn[4], chain, icode, het;
        int resid, model, part;
    };
    std::vector<Atom> atoms;
    atoms.reserve(capacity > 0 ? capacity : 1024);
    std::unordered_set<std::string> altloc_seen;
    int cur_model = 0, n_models_seen = 0;
    bool in_model = false;
    // chain parts of the current model: order of first appearance of each chain name
    std::vector<char> chain_order;
    auto chain_rank = [&](char c) {
        for (size_t i = 0; i < chain_order.size(); ++i)
            if (chain_order[i] == c) return (int)i;
        chain_order.push_back(c);
        return (int)chain_order.size() - 1;
    };
    // het flag of a residue = record type of its first atom (consecutive atoms with equal chain, number, icode, name)
    char res_het = 'A';
    int res_id = 0, res_model = -1;
    char res_chain = 0, res_icode = 0, res_name[4] = {0, 0, 0, 0};

    for (size_t i = 0; i < len;) {
        const char *e = (const char *)memchr(text + i, '\n', len - i);
        size_t ll = e ? (size_t)(e - (text + i)) : len - i;
        Line l{text + i, (int)std::min<size_t>(ll, 80)};
        i += ll + 1;
        if (l.n > 0 && l.p[l.n - 1] == '\r') --l.n;
        if (starts(l, "MODEL")) {
            if (in_model || n_models_seen > 0 || !atoms.empty()) ++cur_model;
            in_model = true;
            ++n_models_seen;
            chain_order.clear();
            continue;
        }
        if (starts(l, "ENDMDL")) {
            in_model = false;
            continue;
        }
        if (l.n >= 3 && memcmp(l.p, "END", 3) == 0 && (l.n == 3 || isspace((unsigned char)l.p[3]))) break;
        const bool is_atom = starts(l, "ATOM  "), is_het = starts(l, "HETATM");
        if (!is_atom && !is_het) continue;
        if (l.n < 54) {
            set_error("pdb_parse: ATOM/HETATM record shorter than 54 columns");
            return PESTO_EINVAL;
        }
        char f[16];
        Atom a;
        memset(&a, 0, sizeof a);
        field(l, 12, 16, f);
        char raw_name[5];
        memcpy(raw_name, f, 5);
        strip_copy(f, a.name, 4);
        const char altloc = l.p[16];
        field(l, 17, 20, f);
        strip_copy(f, a.resn, 3);
        a.chain = l.n > 21 ? l.p[21] : ' ';
        field(l, 22, 26, f);
        a.resid = atoi(f);
        a.icode = l.n > 26 ? l.p[26] : ' ';
        field(l, 30, 38, f);
        a.x = (float)strtod(f, nullptr);
        field(l, 38, 46, f);
        a.y = (float)strtod(f, nullptr);
        field(l, 46, 54, f);
        a.z = (float)strtod(f, nullptr);
        field(l, 60, 66, f);
        a.b = (float)strtod(f, nullptr);
        field(l, 76, 78, f);
        char el[3];
        strip_copy(f, el, 2);
        el[2] = 0;
        if (!el[0]) {                                   // blank element columns: derive from the atom name columns
            if (isalpha((unsigned char)raw_name[0]) && isalpha((unsigned char)raw_name[1]) && strlen(a.name) < 4) {
                el[0] = raw_name[0];
                el[1] = raw_name[1];
            } else {
                for (int k = 0; k < 4; ++k)
                    if (isalpha((unsigned char)raw_name[k])) {
                        el[0] = raw_name[k];
                        break;
                    }
            }
        }
        a.elem[0] = (char)toupper((unsigned char)el[0]);
        a.elem[1] = el[1] ? (char)tolower((unsigned char)el[1]) : 0;
        a.model = cur_model;
        // residue het flag
        if (!(res_model == a.model && res_chain == a.chain && res_id == a.resid && res_icode == a.icode &&
              memcmp(res_name, a.resn, 4) == 0)) {
            res_model = a.model; res_chain = a.chain; res_id = a.resid; res_icode = a.icode;
            memcpy(res_name, a.resn, 4);
            res_het = is_het ? 'H' : 'A';
        }
        a.het = res_het;
        if (altloc != ' ' && altloc != 0) {
            std::string key;
            key.push_back(a.chain);
            key.push_back('_');
            key += std::to_string(a.resid);
            key.push_back('_');
            key += a.name;
            if (!altloc_seen.insert(key).second) continue;
        }
        a.part = chain_rank(a.chain);
        atoms.push_back(a);
    }
    // chain parts merged per model: stable order by (model, first appearance of the chain name)
    std::stable_sort(atoms.begin(), atoms.end(), [](const Atom &u, const Atom &v) {
        return u.model != v.model ? u.model < v.model : u.part < v.part;
    });
    const int n = (int)atoms.size();
    *n_out = n;
    if (n > capacity) {
        set_error("pdb_parse: %d atoms do not fit the caller's capacity %d", n, capacity);
        return PESTO_EINVAL;
    }
    for (int k = 0; k < n; ++k) {
        const Atom &a = atoms[k];
        xyz[3 * k] = a.x; xyz[3 * k + 1] = a.y; xyz[3 * k + 2] = a.z;
        memcpy(name4 + 4 * k, a.name, 4);
        memcpy(element2 + 2 * k, a.elem, 2);
        memcpy(resname3 + 3 * k, a.resn, 3);
        resid[k] = a.resid;
        het[k] = a.het;
        chain[k] = a.chain;
        model[k] = a.model;
        icode[k] = a.icode == ' ' ? 0 : a.icode;
        bfactor[k] = a.b;
    }
    return PESTO_OK;
}
