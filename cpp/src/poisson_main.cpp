// poisson_main -- the reference's executable (executables/poisson-main.cpp:23-72) on the B200 path, without GDAL, OpenCV or
// spdlog:   poisson_main <input.tif> <replacement.tif> [--reference-layout]
// Bands 1-5 of <input> are blended against bands 1-5 of <replacement> inside the mask made from band 6 of <input> by an
// 11 x 11 morphological close; the result is a copy of <input> with bands 1-5 replaced, written to
// <dir of input>/poisson_simple_replace/<name of input>.  GeoTIFFs through utils/geotiff.h, pixels through the `approx`
// shim over the C-ABI (sa_morph_close_mask, sa_poisson_blend).  There is no CPU solve: without a device it exits 2.
#include <approx/poisson.h>
#include <utils/geotiff.h>

#include <cstdio>
#include <cstring>
#include <memory>

int main(int argc, char** argv)
{
    utils::Layout layout = utils::Layout::Raster;
    std::vector<std::string> args;
    for (int i = 1; i < argc; ++i) {
        if (std::strcmp(argv[i], "--reference-layout") == 0)
            layout = utils::Layout::Reference;
        else
            args.emplace_back(argv[i]);
    }
    if (args.size() != 2) {
        std::fprintf(stderr, "Usage: %s input_path replacement_path [--reference-layout]\n", argv[0]);
        return -1;
    }
    fs::path input(args[0]), replacement(args[1]);
    for (auto const& p : { input, replacement })
        if (!fs::exists(p)) {
            std::fprintf(stderr, "%s does not exist\n", p.string().c_str());
            return -1;
        }
    try {
        std::vector<int> bands = { 1, 2, 3, 4, 5 };
        int cloud_band = 6;
        utils::GeoTIFF<f64> tiff(input.string(), layout);
        auto input_bands = tiff.read(bands);
        MatX<bool> cloudmask;
        try {
            cloudmask = approx::preprocess_cloud_band(tiff.read(cloud_band));
        } catch (std::runtime_error const& e) {
            std::fprintf(stderr, "%s\n", e.what());
            return 2;  // no device (or a CUDA error): there is no CPU fallback
        }
        std::fprintf(stderr, "Finished close + dilate\n");
        utils::GeoTIFF<f64> replacement_tiff(replacement.string(), layout);
        auto replacement_bands = replacement_tiff.read(bands);
        std::fprintf(stderr, "Starting solver...\n");
        auto res = std::make_shared<std::vector<MatX<f64>>>(
            approx::blend_images_poisson(input_bands, replacement_bands, cloudmask));
        std::fprintf(stderr, "Finished solving. Writing results\n");
        utils::GeoTiffWriter<f64> writer(res, input, layout);
        writer.write(input.parent_path() / "poisson_simple_replace" / input.filename());
    } catch (std::exception const& e) {
        std::fprintf(stderr, "poisson_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
