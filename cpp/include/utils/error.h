// utils/error.h -- the exception types of the reference (lib/utils/include/utils/error.h:11-43) that the fill path and its
// callers throw, without spdlog / sqlite3: the message goes to stderr where the reference logs it.
#pragma once

#include <exception>
#include <filesystem>
#include <iostream>
#include <string>
#include <string_view>

namespace fs = std::filesystem;

namespace utils {
class IOError : public std::exception {
public:
    IOError(std::string_view msg, fs::path path) : m_message(msg), m_path(std::move(path))
    {
        std::cerr << "[error] " << m_message << " (path: " << m_path.string() << ")\n";  // error.cpp:7-12
    }
    char const* what() const noexcept override { return m_message.c_str(); }
    fs::path path() const { return m_path; }

private:
    std::string m_message;
    fs::path m_path;
};

class GenericError : public std::exception {
public:
    explicit GenericError(std::string_view msg) : m_message(msg) {}
    char const* what() const noexcept override { return m_message.c_str(); }

private:
    std::string m_message;
};
}  // namespace utils
