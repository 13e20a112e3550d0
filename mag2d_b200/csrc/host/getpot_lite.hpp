// getpot_lite.hpp — the slice of GetPot's interface that mag2d uses, written from its documented behaviour.
//
// The reference reads config.txt and its command line through the third-party GetPot 1.1.18 header
// (reference INSTALL:3, src/param.hpp:4, src/test.cpp:21-34), which is not part of the reference checkout
// and not installed here.  Semantics kept: '#' starts a comment, "key = value" defines a variable,
// "[section]" prefixes the variables that follow with "section/", command-line tokens of the form key=value
// are variables too, numbers are parsed as doubles (so "16e5" is a valid integer) and an unparsable or
// missing value yields the caller's default.
#pragma once
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

class GetPot
{
  public:
    GetPot() = default;
    GetPot(int argc, char** argv)
    {
        for (int a = 1; a < argc; a++) define(argv[a]);
    }
    explicit GetPot(const char* file) { read(file); }
    explicit GetPot(const std::string& file) { read(file.c_str()); }

    void set_prefix(const char* p) { prefix = p; }
    void set(const std::string& key, const std::string& value) { table[key] = value; }
    bool have(const std::string& key) const { return table.count(prefix + key) != 0; }

    double operator()(const char* key, double fallback) const
    {
        auto it = table.find(prefix + key);
        double v;
        if (it == table.end() || std::sscanf(it->second.c_str(), "%lf", &v) != 1) return fallback;
        return v;
    }
    int operator()(const char* key, int fallback) const { return (int)(*this)(key, (double)fallback); }
    std::string operator()(const char* key, const char* fallback) const
    {
        auto it = table.find(prefix + key);
        return it == table.end() ? std::string(fallback) : it->second;
    }

  private:
    std::map<std::string, std::string> table;
    std::string prefix, section;

    static std::string strip(const std::string& s)
    {
        const char* ws = " \t\r\n";
        const size_t b = s.find_first_not_of(ws);
        if (b == std::string::npos) return std::string();
        return s.substr(b, s.find_last_not_of(ws) - b + 1);
    }
    void define(const std::string& token)
    {
        const size_t eq = token.find('=');
        if (eq == std::string::npos || eq == 0) return;
        table[section + strip(token.substr(0, eq))] = strip(token.substr(eq + 1));
    }
    void read(const char* file)
    {
        std::ifstream in(file);
        std::string line;
        while (std::getline(in, line))
        {
            line = strip(line.substr(0, line.find('#')));
            if (line.empty()) continue;
            if (line.front() == '[')
            {
                const std::string name = strip(line.substr(1, line.find(']') == std::string::npos ? std::string::npos : line.find(']') - 1));
                section = name.empty() ? std::string() : name + "/";
                continue;
            }
            const size_t eq = line.find('=');
            if (eq == std::string::npos || eq == 0) continue;
            std::istringstream rest(line.substr(eq + 1));
            std::string first;
            rest >> first;   // the value is the first token after '='
            table[section + strip(line.substr(0, eq))] = first;
        }
    }
};
