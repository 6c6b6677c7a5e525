// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
// Minimal `std::experimental::mdspan` (the subset of the Kokkos reference implementation
// that Basix 0.6 vendors as <basix/mdspan.hpp>) so that the UNCHANGED reference sources
// under /root/reference/cpp compile in this image (no Basix/DOLFINx available).
// Supported: dextents<size_t, N>, layout_right, layout_stride, submdspan with integral,
// std::pair and full_extent slices.  Written from the P0009 interface, not copied.
#pragma once

#include <array>
#include <cassert>
#include <cstdio>
#ifdef EQLB_SHIM_CHECK
#include <execinfo.h>
#endif
#include <cstdlib>
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace std::experimental
{
inline constexpr std::size_t dynamic_extent = static_cast<std::size_t>(-1);

struct full_extent_t
{
  explicit full_extent_t() = default;
};
inline constexpr full_extent_t full_extent{};

template <typename I, std::size_t N>
struct dextents
{
  using index_type = I;
  static constexpr std::size_t rank() { return N; }
  std::array<I, N> e{};
  constexpr I extent(std::size_t i) const { return e[i]; }
};

template <typename I, std::size_t... Es>
struct extents; // static extents are not used by the reference

struct layout_right
{
};
struct layout_stride
{
};

template <typename T, typename Extents, typename Layout = layout_right>
class mdspan;

template <typename T, typename I, std::size_t N, typename Layout>
class mdspan<T, dextents<I, N>, Layout>
{
public:
  using element_type = T;
  using value_type = std::remove_cv_t<T>;
  using index_type = I;
  using size_type = std::size_t;
  using extents_type = dextents<I, N>;
  using layout_type = Layout;
  using reference = T&;
  using data_handle_type = T*;

  constexpr mdspan() = default;

  template <typename... Ix, typename = std::enable_if_t<sizeof...(Ix) == N && (std::is_convertible_v<Ix, I> && ...)>>
  constexpr mdspan(T* p, Ix... ext) : _p(p), _e{static_cast<I>(ext)...}
  {
    set_right_strides();
  }

  template <typename J>
  constexpr mdspan(T* p, const std::array<J, N>& ext) : _p(p)
  {
    for (std::size_t i = 0; i < N; ++i)
      _e[i] = static_cast<I>(ext[i]);
    set_right_strides();
  }

  constexpr mdspan(T* p, const extents_type& ext) : _p(p), _e(ext.e) { set_right_strides(); }

  // raw constructor used by submdspan
  constexpr mdspan(T* p, const std::array<I, N>& ext, const std::array<I, N>& str, int) : _p(p), _e(ext), _s(str) {}

  // converting constructor: non-const -> const, layout_right -> layout_stride, same -> same
  template <typename U, typename L2,
            typename = std::enable_if_t<std::is_convertible_v<U (*)[], T (*)[]>
                                        && (std::is_same_v<L2, Layout> || std::is_same_v<Layout, layout_stride>)>>
  constexpr mdspan(const mdspan<U, dextents<I, N>, L2>& o) : _p(o.data_handle()), _e(o.extent_array()), _s(o.stride_array())
  {
  }

  static constexpr std::size_t rank() { return N; }
  constexpr I extent(std::size_t i) const { return _e[i]; }
  constexpr I stride(std::size_t i) const { return _s[i]; }
  constexpr extents_type extents() const { return extents_type{_e}; }
  constexpr T* data_handle() const { return _p; }
  constexpr std::size_t size() const
  {
    std::size_t n = 1;
    for (std::size_t i = 0; i < N; ++i)
      n *= _e[i];
    return n;
  }
  constexpr bool empty() const { return size() == 0; }
  constexpr const std::array<I, N>& extent_array() const { return _e; }
  constexpr const std::array<I, N>& stride_array() const { return _s; }

  template <typename... Ix, typename = std::enable_if_t<sizeof...(Ix) == N>>
  constexpr T& operator()(Ix... idx) const
  {
#ifdef EQLB_SHIM_CHECK
    {
      const std::array<I, N> ix{static_cast<I>(idx)...};
      for (std::size_t i = 0; i < N; ++i)
        if (!(ix[i] < _e[i]))
        {
          std::fprintf(stderr, "mdspan shim: index %zu out of range [0,%zu) in dimension %zu\n", (std::size_t)ix[i], (std::size_t)_e[i], i);
          void* bt[32];
          backtrace_symbols_fd(bt, backtrace(bt, 32), 2);
          std::abort();
        }
    }
#endif
    return _p[offset(std::index_sequence_for<Ix...>{}, idx...)];
  }

private:
  template <std::size_t... Is, typename... Ix>
  constexpr I offset(std::index_sequence<Is...>, Ix... idx) const
  {
    if constexpr (std::is_same_v<Layout, layout_right>)
    {
      // Horner form over the extents: no stride loads, the last index is contiguous
      I off = 0;
      ((off = off * _e[Is] + static_cast<I>(idx)), ...);
      return off;
    }
    else
      return ((static_cast<I>(idx) * _s[Is]) + ... + I(0));
  }
  constexpr void set_right_strides()
  {
    I s = 1;
    for (std::size_t i = N; i-- > 0;)
    {
      _s[i] = s;
      s *= _e[i];
    }
  }
  T* _p = nullptr;
  std::array<I, N> _e{};
  std::array<I, N> _s{};
};

namespace detail
{
template <typename S>
inline constexpr bool is_index_v = std::is_convertible_v<S, std::size_t> && !std::is_same_v<std::decay_t<S>, full_extent_t>;
template <typename S>
inline constexpr bool is_full_v = std::is_same_v<std::decay_t<S>, full_extent_t>;

template <typename S>
struct is_pair : std::false_type
{
};
template <typename A, typename B>
struct is_pair<std::pair<A, B>> : std::true_type
{
};
template <typename A, typename B>
struct is_pair<std::tuple<A, B>> : std::true_type
{
};

// layout_right is preserved iff the slices read [index]* [range|full]? [full]*
template <typename... S>
constexpr bool keeps_layout_right()
{
  constexpr bool idx[] = {is_index_v<std::decay_t<S>>...};
  constexpr bool full[] = {is_full_v<S>...};
  std::size_t i = 0;
  const std::size_t n = sizeof...(S);
  while (i < n && idx[i])
    ++i;
  if (i < n && !idx[i])
    ++i; // one range or full
  while (i < n && full[i])
    ++i;
  return i == n;
}

template <typename I>
struct slice_t
{
  I lo, len;
  bool keep;
};

template <typename I, typename S>
constexpr slice_t<I> make_slice(const S& s, I ext)
{
  using D = std::decay_t<S>;
  if constexpr (is_full_v<D>)
    return {0, ext, true};
  else if constexpr (is_pair<D>::value)
    return {static_cast<I>(std::get<0>(s)), static_cast<I>(std::get<1>(s)) - static_cast<I>(std::get<0>(s)), true};
  else
    return {static_cast<I>(s), 1, false};
}
} // namespace detail

template <typename T, typename I, std::size_t N, typename Layout, typename... Slices>
constexpr auto submdspan(const mdspan<T, dextents<I, N>, Layout>& src, Slices... slices)
{
  static_assert(sizeof...(Slices) == N);
  constexpr std::size_t R = ((detail::is_index_v<std::decay_t<Slices>> ? 0 : 1) + ...);
  using out_layout
      = std::conditional_t<std::is_same_v<Layout, layout_right> && detail::keeps_layout_right<Slices...>(), layout_right, layout_stride>;
  std::array<detail::slice_t<I>, N> sl{};
  {
    std::size_t d = 0;
    ((sl[d] = detail::make_slice<I>(slices, src.extent(d)), ++d), ...);
  }
  I off = 0;
  std::array<I, R> e{}, s{};
  std::size_t r = 0;
  for (std::size_t d = 0; d < N; ++d)
  {
    off += sl[d].lo * src.stride(d);
    if (sl[d].keep)
    {
      e[r] = sl[d].len;
      s[r] = src.stride(d);
      ++r;
    }
  }
  return mdspan<T, dextents<I, R>, out_layout>(src.data_handle() + off, e, s, 0);
}
} // namespace std::experimental
