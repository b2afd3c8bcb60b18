// Multi-record FASTA front-end of the batch search (row "next" #4 of the path: the reference reads the file
// with Fasta::load, cppsrc/fa.cpp:22-60, and then searches only its FIRST record, cppsrc/biorseo.cpp:144-147;
// the batch search takes them all).
#ifndef BIORSEO_B200_FASTA_H_
#define BIORSEO_B200_FASTA_H_

#include <string>
#include <vector>

namespace bsh {

struct FastaRecord {
    std::string name;   // header line without '>'
    std::string seq;    // sequence lines, each cut at its first non-letter
    std::string str;    // optional structure lines (characters of "()[].?xle ")
};

// Same record boundaries and line rules as the reference's loader: a line starting with '>' opens a record (the
// previous one is kept only if its name is not empty); a line whose first character is one of "()[].?xle " — or
// an empty line — is a structure line, kept up to its first character outside that set; any other line is a
// sequence line, kept up to its first non-letter. Lines before the first header are lost (the header clears them).
// Returns false (and `err`) when the file cannot be opened, or when a record that is followed by another one has a
// structure whose length differs from its sequence's (the reference aborts on an assertion there).
bool load_fasta(const std::string& path, std::vector<FastaRecord>& records, std::string& err);

}   // namespace bsh
#endif
