// utils/filesystem.h -- utils::find_directory_contents (lib/utils/include/utils/filesystem.h:7-13,
// lib/utils/source/filesystem.cpp:3-15) on std::regex instead of Boost.Regex.
#pragma once

#include <filesystem>
#include <regex>

namespace fs = std::filesystem;

namespace utils {
enum DirectoryContents { NoSatelliteData, MultiSpectral, Radar };

inline DirectoryContents find_directory_contents(fs::path const& path)
{
    static const std::regex expr { R"(\d{4}-\d{2}-\d{2})" };
    if (!std::regex_match(path.filename().string(), expr))
        return DirectoryContents::NoSatelliteData;
    return fs::exists(path / fs::path("B04.tif")) ? DirectoryContents::MultiSpectral : DirectoryContents::Radar;
}
}  // namespace utils
