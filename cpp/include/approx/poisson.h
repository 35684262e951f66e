// approx/poisson.h -- same declarations as the reference for the mask overloads (lib/approx/include/approx/poisson.h:
// 12-52); implemented in cpp/src/approx_satfill.cpp on top of libsatfill.so.
#pragma once

#include <type_traits>

#include <optional>
#include <string>

#include "utils.h"
#include "utils/date.h"
#include "utils/error.h"

namespace approx {

struct PerfInfo {  // poisson.h:12-21
    long region_size = 0;
    f64 tolerance = 0.0;
    long max_iterations = 0;
    long iterations = 0;
    f64 error = 0.0;
    f64 solve_time = 0.0;
    void write(fs::path const& output) const;  // appends one CSV line (poisson.cpp:14-19)
};

// Offset / white-key overload (poisson.h:30-33, poisson.cpp:21-143): the replacement image is pasted at
// (start_row, start_column); its unknowns are the pixels that are NOT the white key (MultiChannelImage::valid_pixel).
// The reference solves with a direct factorisation; here the system goes through the same CG (tolerance 1e-12).
void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images, int start_row,
    int start_column);
void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images,
    MatX<bool> const& invalid_mask, f64 tolerance = 1e-6, std::optional<int> max_iterations = {});
std::vector<MatX<f64>> blend_images_poisson(std::vector<MatX<f64>> const& input_images,
    std::vector<MatX<f64>> const& replacement_images, MatX<bool> const& invalid_mask, f64 tolerance = 1e-6,
    std::optional<int> max_iterations = {});

// poisson.h:54, poisson.cpp:305-321: paint the pasted (non white-key) pixels of the replacement with `color` in the first
// three channels of the input, at the offset.
void highlight_area_replaced(MultiChannelImage& input_images, MultiChannelImage const& replacement_images, int start_row,
    int start_col, Vec3<f64> const& color);

// approx::DayInfo (lib/approx/include/approx/db.h:12-17) and the ranking rule of find_good_close_image
// (poisson.cpp:323-349) on a list of candidates.  The reference pulls the candidates out of SQLite (approx::DataBase,
// db.cpp:97-156); that side needs SQLiteCpp + Boost.date_time, which this image lacks, so the C++ shim takes the rows as
// arguments and the database itself is mirrored in the Python package (satellite_approximation_b200/scenes.py).
struct DayInfo {
    utils::Date date;
    f64 percent_invalid = 0.0;
    [[nodiscard]] f64 distance(utils::Date const& other, f64 weight) const  // db.cpp:12-16
    {
        return weight * (f64)std::labs(other.days() - date.days()) + (1 - weight) * percent_invalid;
    }
};
// Returns the ISO date of the best neighbour, `date_string` itself when the date has fewer invalid pixels than that
// neighbour, "" when `close_images` is empty; throws utils::GenericError for a weight outside [0, 1].
std::string find_good_close_image(std::string const& date_string, f64 distance_weight, std::vector<DayInfo> close_images,
    f64 percent_invalid_of_date);

// The reference's own signature (lib/approx/include/approx/poisson.h:63, poisson.cpp:323-349):
//     std::string find_good_close_image(std::string const& date_string, f64 distance_weight, DataBase& db);
// as a template over the database type, so that the header needs neither SQLiteCpp nor Boost: any type with the two
// queries of approx::DataBase (db.h:30-31) fits -- the reference's own class compiled from its own db.cpp, or a stand-in
// (tests).  The rows' `date` is a utils::Date or anything with year() / month() / day() (boost::gregorian::date).  Same
// order of effects as upstream: the weight is checked before the database is touched, and the second query only runs
// when a neighbour exists.
namespace detail {
template <class D>
utils::Date to_date(D const& d)
{
    if constexpr (std::is_same_v<std::decay_t<D>, utils::Date>)
        return d;
    else
        return utils::Date((int)d.year(), (int)d.month(), (int)d.day());
}
}  // namespace detail
template <class DataBaseT>
std::string find_good_close_image(std::string const& date_string, f64 distance_weight, DataBaseT& db)
{
    if (distance_weight < 0 || distance_weight > 1)  // poisson.cpp:325-327
        throw utils::GenericError("Could not find close image: distance weight not between 0 and 1");
    std::vector<DayInfo> rows;
    for (auto const& r : db.select_close_images(date_string))
        rows.push_back(DayInfo { detail::to_date(r.date), (f64)r.percent_invalid });
    if (rows.empty())  // poisson.cpp:331-334
        return {};
    return find_good_close_image(date_string, distance_weight, std::move(rows),
        (f64)db.select_info_about_date(date_string).percent_invalid);
}

// preprocess_cloud_band of poisson_main (executables/poisson-main.cpp:10-21): (2 dilation_size + 1)^2 rectangular
// morphological close of the cloud band, cast to bool -- on the GPU (sa_morph_close_mask), bit-exact against
// cv::morphologyEx.  Lives in the executable upstream; here so that a poisson_main can be linked without OpenCV.
MatX<bool> preprocess_cloud_band(MatX<f64> const& cloud_band, int dilation_size = 5);

// The record of the last blend (the reference appends it to a hard-coded path, poisson.cpp:287-289).
PerfInfo const& last_perf_info();

}  // namespace approx
