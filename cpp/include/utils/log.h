// utils/log.h -- the logging surface of the reference's lib/utils (source/log.cpp:8-37: named spdlog loggers, console sink at
// `warn`, file sink at `trace` under ./logs/) reduced to what the fill path uses: messages at spdlog's levels, written to
// stderr when at or above the current level.  Importing the module or linking the shim creates no directories and no files
// (SURVEY.md App. B7).  Level values are spdlog's (trace 0 .. critical 5, off 6); src/main.cpp:24-34 binds Debug..Critical.
#pragma once

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace utils {

enum class LogLevel : int { trace = 0, debug = 1, info = 2, warn = 3, err = 4, critical = 5, off = 6 };

inline std::atomic<int>& log_level_ref()
{
    static std::atomic<int> level { (int)LogLevel::warn };  // the reference's console sink level (log.cpp:14)
    return level;
}
inline void set_log_level(LogLevel level) { log_level_ref().store((int)level); }
inline LogLevel log_level() { return (LogLevel)log_level_ref().load(); }

#if defined(__GNUC__)
__attribute__((format(printf, 3, 4)))
#endif
inline void log(LogLevel level, const char* logger, const char* fmt, ...)
{
    if ((int)level < log_level_ref().load() || level == LogLevel::off)
        return;
    static const char* names[] = { "trace", "debug", "info", "warning", "error", "critical" };
    std::fprintf(stderr, "[%s] [%s] ", logger, names[(int)level]);
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(stderr, fmt, ap);
    va_end(ap);
    std::fputc('\n', stderr);
}

}  // namespace utils
