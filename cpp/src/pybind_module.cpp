// satellite_approximation._core -- the fill-path functions of the reference's pybind11 module (src/main.cpp:16-58)
// bound to the B200 implementation.  Same names, argument names, noconvert rules and defaults.
#include <pybind11/eigen.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <filesystem>
#include <string>

#include <approx/laplace.h>
#include <approx/poisson.h>
#include <utils/filesystem.h>
#include <utils/geotiff.h>
#include <utils/log.h>

#include <algorithm>

#include <array>

namespace py = pybind11;
using namespace py::literals;

enum class PyLogLevel { Debug = 1, Info = 2, Warn = 3, Error = 4, Critical = 5 };  // spdlog values (src/main.cpp:24-29)

PYBIND11_MODULE(_core, m)
{
    m.doc() = "Data processing for sentinel satellite imagery (Laplace / Poisson fill path on B200)";
    py::class_<std::filesystem::path>(m, "Path")  // src/main.cpp:20-22
        .def(py::init<std::string>())
        .def("__str__", [](std::filesystem::path const& p) { return p.string(); })
        .def("__fspath__", [](std::filesystem::path const& p) { return p.string(); });
    py::implicitly_convertible<std::string, std::filesystem::path>();
    py::enum_<PyLogLevel>(m, "LogLevel")
        .value("Debug", PyLogLevel::Debug)
        .value("Info", PyLogLevel::Info)
        .value("Warn", PyLogLevel::Warn)
        .value("Error", PyLogLevel::Error)
        .value("Critical", PyLogLevel::Critical);
    // src/main.cpp:30-34 sets the level of the reference's spdlog loggers; here: of the shim's stderr log (utils/log.h)
    m.def("set_log_level", [](PyLogLevel level) { utils::set_log_level((utils::LogLevel)(int)level); });
    // the record of the last fill (approx::PerfInfo, poisson.h:12-21; the reference appends it to a hard-coded CSV)
    m.def("last_perf_info", []() {
        approx::PerfInfo const& p = approx::last_perf_info();
        return py::dict("region_size"_a = p.region_size, "tolerance"_a = p.tolerance, "max_iterations"_a = p.max_iterations,
            "iterations"_a = p.iterations, "error"_a = p.error, "solve_time"_a = p.solve_time);
    });
    // apply_laplace on numpy arrays (laplace.h:31 takes cv::Mat: uint8 H x W x 3 in cv::imread order)
    m.def(
        "apply_laplace",
        [](py::array_t<unsigned char, py::array::c_style | py::array::forcecast> image,
            py::array_t<unsigned char, py::array::c_style | py::array::forcecast> invalid_image, double red_threshold) {
            if (image.ndim() != 3 || image.shape(2) != 3 || invalid_image.ndim() != 3 || invalid_image.shape(2) != 3)
                throw std::runtime_error("apply_laplace: uint8 H x W x 3 images");
            if (image.shape(0) != invalid_image.shape(0) || image.shape(1) != invalid_image.shape(1))
                throw std::runtime_error("Input image and mask need to be the same size");  // laplace.cpp:124-127
            py::array_t<double> out({ image.shape(0), image.shape(1), (py::ssize_t)3 });
            {
                py::gil_scoped_release release;
                approx::apply_laplace(image.data(), invalid_image.data(), image.shape(0), image.shape(1), red_threshold,
                    out.mutable_data());
            }
            return out;
        },
        "image"_a, "invalid_image"_a, "red_threshold"_a = 220.0);
    m.def("get_log_level", []() { return (PyLogLevel)std::min(std::max((int)utils::log_level(), 1), 5); });
    m.def(
        "filling_missing_portions_smooth_boundaries",
        [](MatX<f64>& input_image, MatX<bool> const& invalid_pixels) {
            py::gil_scoped_release release;
            approx::fill_missing_portion_smooth_boundary(input_image, invalid_pixels);
            return std::move(input_image);  // the caster's own copy of the caller's array: hand it back without another copy
        },
        py::arg("input_image").noconvert(), py::arg("invalid_pixels").noconvert());
    m.def("blend_images_poisson",
        py::overload_cast<std::vector<MatX<f64>> const&, std::vector<MatX<f64>> const&, MatX<bool> const&, f64,
            std::optional<int>>(&approx::blend_images_poisson),
        "input_image"_a, "replacement_image"_a, "invalid_mask"_a, "tolerance"_a = 1e-6, "max_iterations"_a = std::nullopt,
        py::call_guard<py::gil_scoped_release>());
    // not bound by the reference (its offset overload is only reachable from C++, poisson.h:30-33): exposed so that the
    // parity tests can drive the C++ shim's implementation of it
    m.def(
        "blend_images_poisson_offset",
        [](std::vector<MatX<f64>> input_images, std::vector<MatX<f64>> const& replacement_images, int start_row,
            int start_column) {
            py::gil_scoped_release release;
            approx::MultiChannelImage input(std::move(input_images)), replacement(replacement_images);
            approx::blend_images_poisson(input, replacement, start_row, start_column);
            return input.images;
        },
        "input_image"_a, "replacement_image"_a, "start_row"_a, "start_column"_a);
    // the dependency-free host pieces of the shim, exposed for the tests (tests/test_host_api.py)
    m.def(
        "highlight_area_replaced",
        [](std::vector<MatX<f64>> input_images, std::vector<MatX<f64>> const& replacement_images, int start_row,
            int start_column, std::array<double, 3> color) {
            approx::MultiChannelImage input(std::move(input_images)), replacement(replacement_images);
            approx::highlight_area_replaced(input, replacement, start_row, start_column,
                Vec3<f64>(color[0], color[1], color[2]));
            return input.images;
        },
        "input_image"_a, "replacement_image"_a, "start_row"_a, "start_column"_a, "color"_a);
    m.def(
        "find_good_close_image",
        [](std::string const& date_string, double weight, std::vector<std::pair<std::string, double>> const& close,
            double percent_invalid_of_date) {
            std::vector<approx::DayInfo> info;
            for (auto const& [d, p] : close)
                info.push_back({ utils::Date(d), p });
            return approx::find_good_close_image(date_string, weight, std::move(info), percent_invalid_of_date);
        },
        "date_string"_a, "distance_weight"_a, "close_images"_a, "percent_invalid_of_date"_a);
    // utils/geotiff.h (the C++ twin of satellite_approximation_b200/geotiff.py), exposed for tests/test_geotiff.py
    m.def(
        "geotiff_read",
        [](std::string const& path, int band, bool reference_layout) {
            return utils::GeoTIFF<f64>(path, reference_layout ? utils::Layout::Reference : utils::Layout::Raster).read(band);
        },
        "path"_a, "band"_a, "reference_layout"_a = false);
    m.def("geotiff_read_u8", [](std::string const& path, int band) { return utils::GeoTIFF<utils::u8>(path).read(band); });
    m.def("geotiff_read_i16", [](std::string const& path, int band) { return utils::GeoTIFF<utils::i16>(path).read(band); });
    m.def("geotiff_info", [](std::string const& path) {
        utils::GeoTIFF<f64> t(path);
        return py::make_tuple(t.height, t.width, t.raster_count(),
            std::vector<double>(t.geoTransform, t.geoTransform + 6));
    });
    m.def(
        "geotiff_write",
        [](std::vector<MatX<f64>> values, std::string const& template_path, std::string const& destination, int start_index,
            bool reference_layout, bool single) {
            auto layout = reference_layout ? utils::Layout::Reference : utils::Layout::Raster;
            if (single) {
                utils::GeoTiffWriter<f64>(std::make_shared<MatX<f64>>(values.at(0)), template_path, layout)
                    .write(destination, start_index);
            } else {
                utils::GeoTiffWriter<f64>(std::make_shared<std::vector<MatX<f64>>>(std::move(values)), template_path, layout)
                    .write(destination, start_index);
            }
        },
        "values"_a, "template_path"_a, "destination"_a, "start_index"_a = 1, "reference_layout"_a = false, "single"_a = false);
    py::register_exception<utils::IOError>(m, "IOError", PyExc_OSError);
    m.def("find_directory_contents", [](std::string const& path) { return (int)utils::find_directory_contents(path); });
    m.def("date_days", [](std::string const& d) { return utils::Date(d).days(); });
    py::register_exception<utils::GenericError>(m, "GenericError", PyExc_RuntimeError);
    m.def(
        "find_connected_components",
        [](MatX<bool> const& invalid) {
            approx::ConnectedComponents cc = approx::find_connected_components(invalid);
            return py::make_tuple(cc.matrix, cc.region_map.size());
        },
        py::arg("invalid_pixels").noconvert());
    m.def("set_laplace_options", [](double tolerance, long max_iterations, bool multigrid) {
        approx::set_laplace_options({ tolerance, max_iterations, multigrid });
    }, "tolerance"_a = 0.0, "max_iterations"_a = 0, "multigrid"_a = true);
}
