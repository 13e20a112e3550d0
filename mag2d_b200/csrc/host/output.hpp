// output.hpp — output directory + config backups (drop-in for reference src/output.cpp, boost::filesystem
// replaced by std::filesystem).  The interactive "already exists, use it anyway? y/n" prompt is kept.
#pragma once
#include <filesystem>
#include <iostream>
#include <string>

class t_output
{
    std::string output_dir;

  public:
    explicit t_output(const std::string& dir) : output_dir(dir)
    {
        namespace fs = std::filesystem;
        const fs::path path(dir);
        if (fs::exists(path))
        {
            std::cout << path.filename().string() << " already exists, use it anyway? y/n\n";
            std::string response;
            std::cin >> response;
            if (response != "y") exit(1);
            if (!fs::is_directory(path))
            {
                std::cerr << path.filename().string() << " is not a directory\n";
                exit(1);
            }
        }
        else
        {
            std::cout << path << std::endl;
            fs::create_directories(path);
        }
    }
    void backup(const std::string& oldname, const std::string& newname)
    {
        namespace fs = std::filesystem;
        const std::string target = output_dir + "/" + newname;
        std::cout << "making backup of " + oldname + " to " + target << std::endl;
        if (fs::exists(target) && !fs::is_directory(target)) fs::remove(target);
        fs::copy_file(oldname, target);
    }
};
