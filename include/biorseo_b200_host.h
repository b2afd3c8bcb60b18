/*
 * biorseo_b200_host.h — C entry points of the host-side glue (C++: biorseo_b200/host/).
 *
 * These are conveniences for callers that are not C++ (the Python test-suite and bench.py):
 * they expose the motif-library loaders (JSON / CaRNAval RIN / .desc -> flat component
 * table) and the record -> Motif reconstruction that the C++ drop-in
 * (bsh::InsertionSiteSearch, biorseo_b200/host/site_search.h) uses internally. They are
 * not part of the search boundary; that is include/biorseo_b200.h.
 */
#ifndef BIORSEO_B200_HOST_H_
#define BIORSEO_B200_HOST_H_

#include "biorseo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BSH_KIND_JSON 0   /* --jsonfolder : cppsrc/MOIP.cpp:244-296 */
#define BSH_KIND_RIN  1   /* --rinfolder  : cppsrc/MOIP.cpp:199-243 */
#define BSH_KIND_DESC 2   /* --descfolder : cppsrc/MOIP.cpp:146-198 */

typedef struct bsh_library bsh_library;

/* Parse + validate a library from disk. 0 on success. */
int  bsh_library_load(int kind, const char* path, bsh_library** out);
/* Compiled form of a loaded library: one binary file of plain aligned arrays (fixed little-endian header, offset
 * table, sections). bsh_library_load_compiled MAPS the file (mmap): the flat table bsh_library_table returns then
 * points straight into the mapping — no parse, no copy; only module identifiers, link recipes and the reject log
 * are unpacked. Replaces the per-run JSON parse / directory walk / validation of cppsrc/MOIP.cpp:159-188, 211-234,
 * 259-288 and the per-placement RIN re-read of cppsrc/Motif.cpp:115-151. Damaged files are refused. */
int  bsh_library_save_compiled(const bsh_library* lib, const char* path);
int  bsh_library_load_compiled(const char* path, bsh_library** out);
uint64_t bsh_library_mapped_bytes(const bsh_library* lib);                  /* size of the mapping behind the table, 0 if none */
int  bsh_library_kind(const bsh_library* lib);                              /* BSH_KIND_* */
/* Build a JSON-style library in memory (synthetic benchmarks): new, add..., seal. */
int  bsh_library_new(int kind, bsh_library** out);
int  bsh_library_add_json(bsh_library* lib, const char* key, const char* sequence, const char* struct2d); /* reject code or 0 */
int  bsh_library_seal(bsh_library* lib);
void bsh_library_free(bsh_library* lib);

int         bsh_library_table(const bsh_library* lib, bss_table* out);   /* pointers stay owned by lib */
uint32_t    bsh_library_delay(const bsh_library* lib);
uint32_t    bsh_library_modules(const bsh_library* lib);                  /* accepted modules */
uint32_t    bsh_library_seen(const bsh_library* lib);                     /* accepted + rejected */
uint32_t    bsh_library_rejected(const bsh_library* lib);
int         bsh_library_reject(const bsh_library* lib, uint32_t i, const char** id, char* code);
const char* bsh_library_module_id(const bsh_library* lib, uint32_t module);
uint32_t    bsh_library_module_components(const bsh_library* lib, uint32_t module);

/* Device library straight from a host library. */
int bsh_device_library(const bsh_library* lib, int device, bss_library** out);

/* Search records -> the reference's textual view of each insertion site, one per line:
 * identifier \t first-second ... \t a-b ...   (Motif::get_identifier, comp, links_). The
 * text is malloc'ed; release it with bsh_text_free. */
int  bsh_records_text(const bsh_library* lib, const uint32_t* records, uint64_t n_words, char** text, uint64_t* len);
void bsh_text_free(char* text);

/* Multi-record FASTA reader of the batch front-end (biorseo_b200/host/fasta.h; record and line rules of the
 * reference's Fasta::load, cppsrc/fa.cpp:22-60). One line per record: name \t sequence \t structure. */
int  bsh_fasta_text(const char* path, char** text, uint64_t* len);

const char* bsh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
