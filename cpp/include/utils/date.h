// utils/date.h -- utils::Date (lib/utils/include/utils/date.h:11-27, lib/utils/source/date.cpp) without Boost.date_time
// and SQLiteCpp: the same fields, ordering, stream format and hash; days_from_civil replaces boost::gregorian for the
// day arithmetic of DayInfo::distance (lib/approx/source/db.cpp:12-16).  `bind_sql` is not here (no SQLite in this image);
// the database side lives in the Python package (satellite_approximation_b200/scenes.py).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <ostream>
#include <stdexcept>
#include <string>

namespace utils {
struct Date {
    int year = 0;
    int month = 0;
    int day = 0;

    Date() = default;
    Date(int y, int m, int d) : year(y), month(m), day(d) {}
    // "YYYY-MM-DD" (and "YYYY-M-D", '/' as separator): the numeric forms boost::gregorian::from_simple_string accepts
    explicit Date(std::string const& date_string)
    {
        char s1 = 0, s2 = 0;
        int n = 0;
        if (std::sscanf(date_string.c_str(), "%d%c%d%c%d%n", &year, &s1, &month, &s2, &day, &n) != 5
            || (s1 != '-' && s1 != '/') || (s2 != '-' && s2 != '/') || n != (int)date_string.size() || !valid())
            throw std::out_of_range("not a year-month-day date: " + date_string);
    }

    [[nodiscard]] bool valid() const
    {
        static const int len[] = { 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 };
        if (month < 1 || month > 12 || day < 1)
            return false;
        bool leap = (year % 4 == 0 && year % 100 != 0) || year % 400 == 0;
        return day <= len[month - 1] + (month == 2 && leap ? 1 : 0);
    }

    // days since 1970-01-01 in the proleptic Gregorian calendar
    [[nodiscard]] long days() const
    {
        long y = year - (month <= 2);
        long era = (y >= 0 ? y : y - 399) / 400;
        long yoe = y - era * 400;
        long doy = (153 * (month + (month > 2 ? -3 : 9)) + 2) / 5 + day - 1;
        long doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
        return era * 146097 + doe - 719468;
    }

    bool operator==(Date const& o) const { return year == o.year && month == o.month && day == o.day; }
    bool operator<(Date const& o) const { return days() < o.days(); }
    friend std::ostream& operator<<(std::ostream& os, Date const& d)  // date.cpp:33-36
    {
        return os << d.year << '-' << std::setw(2) << std::setfill('0') << d.month << '-' << std::setw(2)
                  << std::setfill('0') << d.day;
    }
};
}  // namespace utils

namespace std {
template <>
struct hash<utils::Date> {
    size_t operator()(utils::Date const& d) const noexcept
    {
        size_t seed = 0;  // boost::hash_combine (date.h:33-39)
        for (int v : { d.year, d.month, d.day })
            seed ^= std::hash<int> {}(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
}  // namespace std
