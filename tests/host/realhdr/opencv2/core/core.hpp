// Test infrastructure (tests/test_host_dropin.py::test_host_sources_parse_against_the_reference_headers): lets the reference's REAL
// headers (include/*.h, Thirdparty/DBoW2) be parsed without OpenCV.  The arithmetic shim is oracle/shim (included below); this file only
// adds declarations of the cv:: persistence types that DBoW2's TemplatedVocabulary.h names (never called on the hot path).
#pragma once
#include_next <opencv2/core/core.hpp>
#include <sstream>
#include <string>
namespace cv {
struct FileNode {
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
}
