// Host-side builder of the flat device library: bss_table -> unit descriptors (bss_layout.h).
// What it replaces (reference, CPU): the per-run library walk + per-module job construction of
// cppsrc/MOIP.cpp:159-188, 211-234, 259-288 — here done once, into arrays the kernels read.
#include "bss_layout.h"

#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "biorseo_b200.h"

namespace bss {

namespace {

bool compatible(uint8_t a, uint8_t b) { return a == '.' || b == '.' || a == b; }

// Can two matches of this pattern overlap, i.e. is the pattern compatible with itself at
// some shift 1..k-1 ?
bool can_self_overlap(const uint8_t* p, uint32_t k)
{
    for (uint32_t s = 1; s < k; s++) {
        bool ok = true;
        for (uint32_t j = 0; ok && j + s < k; j++) ok = compatible(p[j], p[j + s]);
        if (ok) return true;
    }
    return false;
}

// Match program of one pattern: one (occurrence row, shift) step per literal symbol, or per pair of adjacent
// literal symbols when the alphabet is small enough for pair rows.
void append_program(const uint8_t* p, uint32_t k, const int* code_of, bool pairs, uint32_t ns, std::vector<uint16_t>& program)
{
    for (uint32_t j = 0; j < k;) {
        if (p[j] == '.') { j++; continue; }
        uint32_t row = (uint32_t)code_of[p[j]], shift = j;
        if (pairs && j + 1 < k && p[j + 1] != '.') {
            row = ns + row * ns + (uint32_t)code_of[p[j + 1]];
            j += 2;
        } else {
            j += 1;
        }
        program.push_back((uint16_t)(row | ((shift & 31u) << 7) | ((shift >> 5) << 12)));
    }
}

int fail(std::string& error, int code, const char* what)
{
    error = what;
    return code;
}

}   // namespace

int build_library(const bss_table* t, BuiltLibrary& out, std::string& error)
{
    if (!t) return fail(error, BSS_ERR_INVALID, "null argument");
    const uint32_t n_all = t->n_units + t->n_gates;
    if (n_all && (!t->unit_module || !t->unit_delay || !t->unit_comp_begin || !t->comp_pat_begin || !t->unit_check_begin))
        return fail(error, BSS_ERR_INVALID, "null table array");
    if (t->n_units && !t->unit_gate) return fail(error, BSS_ERR_INVALID, "null unit_gate");

    // alphabet: distinct literal pattern bytes
    int code_of[256];
    std::fill(code_of, code_of + 256, -1);
    std::vector<uint8_t>& alphabet = out.alphabet;
    alphabet.clear();
    const uint32_t n_comps = n_all ? t->unit_comp_begin[n_all] : 0;
    const uint32_t n_pat   = n_comps ? t->comp_pat_begin[n_comps] : 0;
    if (n_pat && !t->pattern) return fail(error, BSS_ERR_INVALID, "null pattern array");
    for (uint32_t i = 0; i < n_pat; i++) {
        const uint8_t b = t->pattern[i];
        if (b == '.' || code_of[b] >= 0) continue;
        if (b == '\n' || b == '\r') return fail(error, BSS_ERR_INVALID, "pattern contains a line terminator");
        code_of[b] = (int)alphabet.size();
        alphabet.push_back(b);
    }
    if (alphabet.size() > BSS_MAX_ALPHABET) return fail(error, BSS_ERR_LIMIT, "more than BSS_MAX_ALPHABET distinct pattern bytes");

    // unit descriptors
    std::vector<uint32_t>& blob = out.blob;
    blob.clear();
    out.unit_off.assign(n_all + 1, 0);
    out.module_components.assign(t->n_modules, 0);
    out.unit_probe.assign(4 * (size_t)std::max<uint32_t>(1, t->n_units), 0xFFFFu);   // k-mer probes of up to four components
    out.unit_flags.assign(std::max<uint32_t>(1, t->n_units), 0);
    out.unit_rlen.assign(std::max<uint32_t>(1, t->n_units), 0);
    out.cmax           = 1;
    out.n_window_units = 0;
    std::vector<float> density;   // per searched unit: expected matches per nucleotide, all components
    for (uint32_t u = 0; u < n_all; u++) {
        out.unit_off[u] = (uint32_t)blob.size();
        const uint32_t c0 = t->unit_comp_begin[u], c1 = t->unit_comp_begin[u + 1];
        const uint32_t k0 = t->unit_check_begin[u], k1 = t->unit_check_begin[u + 1];
        if (c1 < c0 || k1 < k0) return fail(error, BSS_ERR_INVALID, "non-monotonic table offsets");
        const uint32_t C = c1 - c0, NK = k1 - k0;
        if (C < 1 || C > BSS_MAX_COMPONENTS) return fail(error, BSS_ERR_LIMIT, "unit with 0 or more than BSS_MAX_COMPONENTS components");
        if (NK > BSS_MAX_CHECKS) return fail(error, BSS_ERR_LIMIT, "unit with more than BSS_MAX_CHECKS checks");
        if (NK && !t->checks) return fail(error, BSS_ERR_INVALID, "null checks array");
        if (t->unit_delay[u] > 0xFFFFu) return fail(error, BSS_ERR_LIMIT, "delay above 65535");
        if (t->unit_module[u] >= t->n_modules) return fail(error, BSS_ERR_INVALID, "unit_module out of range");
        uint32_t gate = kNoGate;
        if (u < t->n_units) {
            gate = t->unit_gate[u];
            if (gate != kNoGate && (gate < t->n_units || gate >= n_all)) return fail(error, BSS_ERR_INVALID, "unit_gate out of range");
            uint32_t& mc = out.module_components[t->unit_module[u]];
            if (mc != 0 && mc != C) return fail(error, BSS_ERR_INVALID, "units of one module disagree on the component count");
            mc               = C;
            out.unit_rlen[u] = (uint8_t)(C + 1);
        }
        out.cmax = std::max(out.cmax, C);

        // order checks by the level at which both ends are known
        std::vector<uint32_t> order(NK);
        std::vector<int>      level(NK);
        for (uint32_t i = 0; i < NK; i++) {
            const int16_t* c = t->checks + 4 * (size_t)(k0 + i);
            if (c[0] < -1 || c[0] >= (int)C || c[2] < -1 || c[2] >= (int)C) return fail(error, BSS_ERR_INVALID, "check anchor out of range");
            level[i] = std::max(0, std::max((int)c[0], (int)c[2]));
            order[i] = i;
        }
        // (within a level the row checks come first: the window walk decodes them once per level entry)
        auto is_row = [&](uint32_t i) {
            const int16_t* c = t->checks + 4 * (size_t)(k0 + i);
            const int      lv = level[i];
            const bool     a_on = c[0] == lv, b_on = c[2] == lv;
            if (a_on == b_on || t->unit_delay[u] < 1) return false;
            const int16_t *first = a_on ? c + 2 : c, *second = a_on ? c : c + 2;
            if (first[0] < 0) return false;
            const int ka = (int)(t->comp_pat_begin[c0 + first[0] + 1] - t->comp_pat_begin[c0 + first[0]]);
            return (int)first[1] <= ka - 1 && second[1] >= 0;
        };
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            if (level[a] != level[b]) return level[a] < level[b];
            return is_row(a) && !is_row(b);
        });

        // match programs: one (occurrence row, shift) step per literal symbol, or per pair of
        // adjacent literal symbols when the alphabet is small enough for pair rows
        const bool     pairs = alphabet.size() <= kPairAlphabet;
        const uint32_t ns    = (uint32_t)alphabet.size();
        std::vector<uint16_t>                      program;
        std::vector<std::pair<uint32_t, uint32_t>> prog_range(C);   // first step, number of steps
        for (uint32_t c = c0; c < c1; c++) {
            const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
            if (t->comp_pat_begin[c + 1] < pb || k > BSS_MAX_PATTERN) return fail(error, BSS_ERR_LIMIT, "component pattern longer than BSS_MAX_PATTERN");
            const uint8_t* p = t->pattern + pb;
            const uint32_t first_step = (uint32_t)program.size();
            append_program(p, k, code_of, pairs, ns, program);
            prog_range[c - c0] = {first_step, (uint32_t)program.size() - first_step};
        }
        if (u < t->n_units) {   // uniform symbols: a pattern with l literal symbols matches one position in |alphabet|^l
            double d = 0.0;
            for (uint32_t c = c0; c < c1; c++) {
                const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
                double         f  = 1.0;
                for (uint32_t j = 0; j < k; j++)
                    if (t->pattern[pb + j] != '.') f /= std::max<double>(2.0, (double)alphabet.size());
                d += f;
            }
            density.push_back((float)d);
        }
        if (u < t->n_units && alphabet.size() <= 4) {   // longest window (4..6) of consecutive literal symbols of the (up to four) probed patterns
            uint32_t slot = 0;
            for (uint32_t c = c0; c < c1 && slot < 4; c++) {
                const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
                uint32_t       run = 0, best = 0, best_end = 0;
                for (uint32_t j = 0; j < k && best < kKmerMax; j++) {
                    run = t->pattern[pb + j] == '.' ? 0 : run + 1;
                    if (std::min(run, kKmerMax) > best) { best = std::min(run, kKmerMax); best_end = j + 1; }
                }
                if (best < kKmerMin) continue;
                uint32_t code = 0;
                for (uint32_t i = 0; i < best; i++) code |= (uint32_t)code_of[t->pattern[pb + best_end - best + i]] << (2 * i);
                out.unit_probe[4 * (size_t)u + slot++] = (uint16_t)(code | ((best - kKmerMin) << 12));
                out.has_probes                         = true;
            }
        }
        const uint32_t pat_words = ((uint32_t)program.size() + 1) / 2;
        const uint32_t head      = 4 + 2 * C;
        const size_t   base      = blob.size();
        blob.resize(base + head + pat_words + 2 * (size_t)NK, 0);
        uint32_t* b  = blob.data() + base;
        b[0] = t->unit_module[u];
        b[1] = C | (NK << 8) | (t->unit_delay[u] << 16);
        b[2] = gate;
        b[3] = (head + pat_words) << 16;
        if (!program.empty()) std::memcpy(b + head, program.data(), program.size() * sizeof(uint16_t));
        uint32_t next_check = 0;
        for (uint32_t c = 0; c < C; c++) {
            const uint32_t pb = t->comp_pat_begin[c0 + c], k = t->comp_pat_begin[c0 + c + 1] - pb;
            b[4 + 2 * c] = k | ((head * 2 + prog_range[c].first) << 8) | (can_self_overlap(t->pattern + pb, k) ? kCompOverlap : 0u);
            const uint32_t begin = next_check;
            while (next_check < NK && level[order[next_check]] == (int)c) next_check++;
            b[5 + 2 * c] = begin | (next_check << 8) | (prog_range[c].second << 16);
        }
        uint32_t*             chk = b + head + pat_words;
        std::vector<uint32_t> row_checks(C, 0);   // row checks per level
        for (uint32_t i = 0; i < NK; i++) {
            const int16_t* c = t->checks + 4 * (size_t)(k0 + order[i]);
            const int      lv = level[order[i]];
            // word 0: the end known before level lv is entered; word 1: the end on component lv
            const bool a_on = c[0] == lv, b_on = c[2] == lv;
            const int16_t *first = c, *second = c + 2;
            uint32_t       flags = 0;
            if (a_on != b_on) {
                if (a_on) std::swap(first, second);
            } else {
                flags = kCheckFixed;
            }
            if (!flags && first[0] >= 0 && first[0] < second[0] && t->unit_delay[u] >= 1) {
                const uint32_t ka = t->comp_pat_begin[c0 + first[0] + 1] - t->comp_pat_begin[c0 + first[0]];
                const uint32_t kb = t->comp_pat_begin[c0 + second[0] + 1] - t->comp_pat_begin[c0 + second[0]];
                if (first[1] >= 0 && (uint32_t)first[1] < ka && second[1] >= 0 && (uint32_t)second[1] < kb) flags |= kCheckOrdered;
                if ((int)first[1] <= (int)ka - 1 && second[1] >= 0) {
                    flags |= kCheckRow;
                    row_checks[lv]++;
                }
            }
            if (!(flags & kCheckRow)) b[4 + 2 * lv] |= kCompGeneric;
            chk[2 * i]     = ((uint32_t)(uint8_t)(int8_t)first[0]) | (((uint32_t)(uint16_t)first[1]) << 8) | flags;
            chk[2 * i + 1] = ((uint32_t)(uint8_t)(int8_t)second[0]) | (((uint32_t)(uint16_t)second[1]) << 8);
        }
        if (u < t->n_units) {
            // window walk: every component after the first must be tied to an earlier one by a row check
            bool window = t->unit_delay[u] >= 1 && head + pat_words + 2 * NK <= kBlobStageWords;
            for (uint32_t c = 1; c < C; c++) window = window && row_checks[c] > 0;
            for (uint32_t c = 0; c < C; c++) {
                if (row_checks[c] > 255) window = false;
                b[5 + 2 * c] |= std::min<uint32_t>(row_checks[c], 255u) << 24;
            }
            if (window) {
                b[3] |= kUnitWindow;
                out.unit_flags[u] |= (uint8_t)kUnitWindow;
                out.n_window_units++;
            }
        }
        if (blob.size() > 0x7FFFFFFFu) return fail(error, BSS_ERR_LIMIT, "library too large");
    }
    out.unit_off[n_all] = (uint32_t)blob.size();

    // pattern table: slots for the patterns shared by several components of searched units
    out.pat_blob.clear();
    out.pat_off.assign(1, 0);
    out.pat_tslot.clear();
    out.unit_slots.assign(4 * (size_t)std::max<uint32_t>(1, t->n_units), (uint16_t)kNoSlot);
    out.n_slots = out.n_tslots = out.n_lean_units = 0;
    {
        std::unordered_map<std::string, uint32_t> uses;   // pattern -> components of searched units that carry it
        const uint32_t searched_comps = t->n_units ? t->unit_comp_begin[t->n_units] : 0;
        for (uint32_t c = 0; c < searched_comps; c++)
            uses[std::string(reinterpret_cast<const char*>(t->pattern) + t->comp_pat_begin[c], t->comp_pat_begin[c + 1] - t->comp_pat_begin[c])]++;
        std::unordered_map<std::string, uint32_t> slot_of;
        const bool     pairs = alphabet.size() <= kPairAlphabet;
        const uint32_t ns    = (uint32_t)alphabet.size();
        for (uint32_t u = 0; u < t->n_units; u++) {
            const uint32_t c0 = t->unit_comp_begin[u], C = t->unit_comp_begin[u + 1] - c0;
            bool           every = C <= 4;
            for (uint32_t c = 0; c < C; c++) {
                const uint32_t    pb = t->comp_pat_begin[c0 + c], k = t->comp_pat_begin[c0 + c + 1] - pb;
                const std::string key(reinterpret_cast<const char*>(t->pattern) + pb, k);
                uint32_t          slot = kNoSlot;
                if (uses[key] >= kSlotMinUses) {
                    auto it = slot_of.find(key);
                    if (it != slot_of.end()) {
                        slot = it->second;
                    } else if (out.n_slots < kSlotMax) {
                        slot         = out.n_slots++;
                        slot_of[key] = slot;
                        std::vector<uint16_t> program;
                        append_program(t->pattern + pb, k, code_of, pairs, ns, program);
                        const bool     overlap = can_self_overlap(t->pattern + pb, k);
                        const uint32_t words   = ((uint32_t)program.size() + 1) / 2;
                        const size_t   base    = out.pat_blob.size();
                        out.pat_blob.resize(base + 6 + words, 0);
                        uint32_t* b = out.pat_blob.data() + base;
                        b[1] = 1u;                               // one component, no check, delay 0
                        b[2] = kNoGate;
                        b[3] = (6u + words) << 16;
                        b[4] = k | (12u << 8) | (overlap ? kCompOverlap : 0u);
                        b[5] = (uint32_t)program.size() << 16;
                        if (!program.empty()) std::memcpy(b + 6, program.data(), program.size() * sizeof(uint16_t));
                        out.pat_off.push_back((uint32_t)out.pat_blob.size());
                        out.pat_tslot.push_back(overlap ? (uint16_t)out.n_tslots++ : (uint16_t)kNoSlot);
                    }
                }
                if (c < 4) out.unit_slots[4 * (size_t)u + c] = (uint16_t)slot;
                every = every && slot != kNoSlot;
            }
            if (every && (out.unit_flags[u] & kUnitWindow) && t->unit_gate[u] == kNoGate) {
                out.unit_flags[u] |= (uint8_t)kUnitLean;
                out.n_lean_units++;
            }
        }
    }
    out.list_density    = 0.f;
    if (!density.empty()) {
        const size_t at = (density.size() - 1) * 98 / 100;
        std::nth_element(density.begin(), density.begin() + at, density.end());
        out.list_density = density[at];
    }
    return BSS_OK;
}

}   // namespace bss
