// utils/types.h -- the value-type aliases of the reference (lib/utils/include/utils/types.h:7-34) that the `approx`
// API is written in.  Source-compatible subset: the fmt helper is not part of the fill path.
#pragma once

#include <Eigen/Dense>

#include <cstdint>

namespace utils {
using u8 = std::uint8_t;
using u16 = std::uint16_t;
using u32 = std::uint32_t;
using u64 = std::uint64_t;
using i8 = std::int8_t;
using i16 = std::int16_t;
using i32 = std::int32_t;
using i64 = std::int64_t;
using f32 = float;
using f64 = double;

template <typename T>
using Vec2 = Eigen::Matrix<T, 2, 1>;
template <typename T>
using Vec3 = Eigen::Matrix<T, 3, 1>;
template <typename T>
using VecX = Eigen::Matrix<T, Eigen::Dynamic, 1>;
template <typename T>
using Mat2 = Eigen::Matrix<T, 2, 2>;
template <typename T>
using MatX = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>;  // column-major, like the reference
}  // namespace utils
