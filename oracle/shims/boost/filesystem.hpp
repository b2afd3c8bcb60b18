// Oracle build shim (test infrastructure, NOT product code).
// Boost.Filesystem is absent from this image; the reference only needs the subset that
// std::filesystem provides under the same names (path, exists, recursive_directory_iterator).
#pragma once
#include <filesystem>
#include <fstream>
#include <stack>
#include <algorithm>
namespace boost { namespace filesystem { using namespace std::filesystem; } }
