#include "biorseo_b200_host.h"

#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "fasta.h"
#include "module_library.h"

struct bsh_library {
    bsh::ModuleLibrary lib;
};

namespace {
thread_local std::string g_host_error;
int host_fail(const std::string& what) { g_host_error = what; return -1; }

// Body of an extern "C" entry point: std::filesystem errors, allocation failures and anything else a loader
// throws become a status code + bsh_last_error() instead of crossing the C boundary.
template <typename F>
int host_guarded(F&& body)
{
    try {
        return body();
    } catch (const std::exception& e) {
        return host_fail(std::string("exception: ") + e.what());
    } catch (...) {
        return host_fail("unexpected exception");
    }
}
}   // namespace

extern "C" {

const char* bsh_last_error(void) { return g_host_error.c_str(); }

int bsh_fasta_text(const char* path, char** text, uint64_t* len)
{
    return host_guarded([&]() -> int {
    if (!path || !text || !len) return host_fail("bad argument");
    std::vector<bsh::FastaRecord> records;
    std::string                   err, out;
    if (!bsh::load_fasta(path, records, err)) return host_fail(err);
    for (const bsh::FastaRecord& r : records) out += r.name + "\t" + r.seq + "\t" + r.str + "\n";
    *text = static_cast<char*>(std::malloc(out.size() + 1));
    if (!*text) return host_fail("out of memory");
    std::memcpy(*text, out.c_str(), out.size() + 1);
    *len = out.size();
    return 0;
    });
}

int bsh_library_load(int kind, const char* path, bsh_library** out)
{
    return host_guarded([&]() -> int {
    if (!path || !out || kind < 0 || kind > 2) return host_fail("bad argument");
    *out = nullptr;
    bsh_library* h = new bsh_library();
    std::string  err;
    if (!h->lib.load((bsh::LibraryKind)kind, path, err)) {
        delete h;
        return host_fail(err);
    }
    *out = h;
    return 0;
    });
}

int bsh_library_save_compiled(const bsh_library* lib, const char* path)
{
    return host_guarded([&]() -> int {
    if (!lib || !path) return host_fail("bad argument");
    std::string err;
    if (!lib->lib.save_compiled(path, err)) return host_fail(err);
    return 0;
    });
}

int bsh_library_load_compiled(const char* path, bsh_library** out)
{
    return host_guarded([&]() -> int {
    if (!path || !out) return host_fail("bad argument");
    *out = nullptr;
    bsh_library* h = new bsh_library();
    std::string  err;
    if (!h->lib.load_compiled(path, err)) {
        delete h;
        return host_fail(err);
    }
    *out = h;
    return 0;
    });
}

int bsh_library_kind(const bsh_library* lib) { return lib ? (int)lib->lib.kind() : -1; }

uint64_t bsh_library_mapped_bytes(const bsh_library* lib) { return lib ? (uint64_t)lib->lib.mapped_bytes() : 0; }

int bsh_library_new(int kind, bsh_library** out)
{
    return host_guarded([&]() -> int {
    if (!out || kind != BSH_KIND_JSON) return host_fail("only JSON-style libraries can be built in memory");
    *out = new bsh_library();
    (*out)->lib.clear(bsh::KIND_JSON);
    return 0;
    });
}

int bsh_library_add_json(bsh_library* lib, const char* key, const char* sequence, const char* struct2d)
{
    return host_guarded([&]() -> int {
    if (!lib || !key || !sequence || !struct2d) return host_fail("bad argument");
    return (int)lib->lib.add_json_module(key, sequence, struct2d);
    });
}

int bsh_library_seal(bsh_library* lib)
{
    return host_guarded([&]() -> int {
    if (!lib) return host_fail("bad argument");
    lib->lib.seal();
    return 0;
    });
}

void bsh_library_free(bsh_library* lib) { delete lib; }

int bsh_library_table(const bsh_library* lib, bss_table* out)
{
    return host_guarded([&]() -> int {
    if (!lib || !out) return host_fail("bad argument");
    *out = lib->lib.table();
    return 0;
    });
}

uint32_t bsh_library_delay(const bsh_library* lib) { return lib ? lib->lib.delay() : 0; }
uint32_t bsh_library_modules(const bsh_library* lib) { return lib ? (uint32_t)lib->lib.modules().size() : 0; }
uint32_t bsh_library_seen(const bsh_library* lib) { return lib ? (uint32_t)lib->lib.n_seen() : 0; }
uint32_t bsh_library_rejected(const bsh_library* lib) { return lib ? (uint32_t)lib->lib.rejected().size() : 0; }

int bsh_library_reject(const bsh_library* lib, uint32_t i, const char** id, char* code)
{
    return host_guarded([&]() -> int {
    if (!lib || i >= lib->lib.rejected().size()) return host_fail("bad argument");
    if (id) *id = lib->lib.rejected()[i].id.c_str();
    if (code) *code = lib->lib.rejected()[i].code;
    return 0;
    });
}

const char* bsh_library_module_id(const bsh_library* lib, uint32_t module)
{
    return (lib && module < lib->lib.modules().size()) ? lib->lib.modules()[module].id.c_str() : "";
}

uint32_t bsh_library_module_components(const bsh_library* lib, uint32_t module)
{
    return (lib && module < lib->lib.modules().size()) ? (uint32_t)lib->lib.modules()[module].k.size() : 0;
}

int bsh_device_library(const bsh_library* lib, int device, bss_library** out)
{
    return host_guarded([&]() -> int {
    if (!lib || !out) return host_fail("bad argument");
    const bss_table t = lib->lib.table();
    int rc = bss_library_create(&t, device, out);
    if (rc != BSS_OK) g_host_error = bss_last_error();
    return rc;
    });
}

int bsh_records_text(const bsh_library* lib, const uint32_t* records, uint64_t n_words, char** text, uint64_t* len)
{
    return host_guarded([&]() -> int {
    if (!lib || !text || !len || (n_words && !records)) return host_fail("bad argument");
    std::string    out;
    const uint32_t n_modules = (uint32_t)lib->lib.modules().size();
    for (uint64_t at = 0; at < n_words;) {
        if (records[at] >= n_modules) return host_fail("record names an unknown module");
        const uint32_t words = lib->lib.record_words(records[at]);
        if (at + words > n_words) return host_fail("truncated record");
        const Motif m = lib->lib.rebuild(records + at);
        out += m.get_identifier();
        out += '\t';
        for (size_t j = 0; j < m.comp.size(); j++) {
            if (j) out += ' ';
            out += std::to_string(m.comp[j].pos.first) + "-" + std::to_string(m.comp[j].pos.second);
        }
        out += '\t';
        for (size_t j = 0; j < m.links_.size(); j++) {
            if (j) out += ' ';
            out += std::to_string(m.links_[j].nts.first) + "-" + std::to_string(m.links_[j].nts.second);
        }
        out += '\n';
        at += words;
    }
    *text = (char*)std::malloc(out.size() + 1);
    if (!*text) return host_fail("out of memory");
    std::memcpy(*text, out.data(), out.size());
    (*text)[out.size()] = 0;
    *len = out.size();
    return 0;
    });
}

void bsh_text_free(char* text) { std::free(text); }

}   // extern "C"
