#include "fasta.h"

#include <cctype>
#include <cstring>
#include <fstream>

namespace bsh {

namespace {
const char kStructureChars[] = "()[].?xle ";
// strchr semantics of the reference: the terminating NUL counts as a member, so an empty line is a structure line
bool is_structure_char(char c) { return std::strchr(kStructureChars, c) != nullptr; }
}   // namespace

bool load_fasta(const std::string& path, std::vector<FastaRecord>& records, std::string& err)
{
    std::ifstream in(path);
    if (!in) {
        err = "cannot open " + path;
        return false;
    }
    FastaRecord cur;
    std::string line;
    while (std::getline(in, line)) {
        const char head = line.empty() ? '\0' : line[0];
        if (head == '>') {
            if (!cur.name.empty()) {
                // the reference asserts this when it closes a record at the next header (cppsrc/fa.cpp:33; its build
                // keeps assertions on) and aborts; the last record of a file is not checked there, nor here
                if (!cur.str.empty() && cur.str.size() != cur.seq.size()) {
                    err = "FASTA record '" + cur.name + "': structure and sequence lengths differ";
                    return false;
                }
                records.push_back(cur);
            }
            cur.name = line.substr(1);
            cur.seq.clear();
            cur.str.clear();
            continue;
        }
        size_t keep = 0;
        if (!is_structure_char(head)) {
            while (keep < line.size() && std::isalpha((unsigned char)line[keep])) keep++;
            cur.seq += line.substr(0, keep);
        } else {
            while (keep < line.size() && is_structure_char(line[keep])) keep++;
            cur.str += line.substr(0, keep);
        }
    }
    if (!cur.name.empty()) records.push_back(cur);
    return true;
}

}   // namespace bsh
