// oracle/shims/boost/filesystem.hpp — TEST INFRASTRUCTURE ONLY.
// boost::filesystem is used by reference src/output.cpp:3-44 (exists, is_directory,
// create_directories, remove, copy_file, path::leaf) and is not installed here; map it onto
// std::filesystem so that the reference's plasma2d / test_MCC mains link unmodified.
#ifndef MAG2D_ORACLE_BOOSTFS_SHIM
#define MAG2D_ORACLE_BOOSTFS_SHIM
#include <filesystem>
#include <ostream>
#include <string>
namespace boost { namespace filesystem {
class path : public std::filesystem::path
{
  public:
    path() {}
    path(const std::string& s) : std::filesystem::path(s) {}
    path(const char* s) : std::filesystem::path(s) {}
    path(const std::filesystem::path& p) : std::filesystem::path(p) {}
    std::string leaf() const { return filename().string(); }
};
inline std::ostream& operator<<(std::ostream& o, const path& p) { return o << '"' << p.string() << '"'; }
inline bool exists(const path& p) { return std::filesystem::exists(p); }
inline bool is_directory(const path& p) { return std::filesystem::is_directory(p); }
inline bool create_directories(const path& p) { return std::filesystem::create_directories(p); }
inline bool remove(const path& p) { return std::filesystem::remove(p); }
inline void copy_file(const path& a, const path& b) { std::filesystem::copy_file(a, b); }
}}
#endif
