// TEST INFRASTRUCTURE: the subset of {fmt} that the reference's
// pair_mtp_extrapolation.cpp uses (fmt::memory_buffer + fmt::format_to with a
// back_inserter, :426-431,:434-441), mapped onto C++20 <format>.
#ifndef LMP_SHIM_FMT_H
#define LMP_SHIM_FMT_H
#include <format>
#include <iterator>
#include <string>
namespace fmt {
class memory_buffer {
  std::string s;

 public:
  using value_type = char;
  void push_back(char c) { s.push_back(c); }
  void clear() { s.clear(); }
  size_t size() const { return s.size(); }
  size_t capacity() const { return s.capacity(); }
  void reserve(size_t n) { s.reserve(n); }
  char *data() { return s.data(); }
  const char *data() const { return s.data(); }
};
template <typename OutputIt, typename... Args>
OutputIt format_to(OutputIt out, const std::string &f, Args &&...args)
{
  return std::vformat_to(out, f, std::make_format_args(args...));
}
template <typename... Args> std::string format(const std::string &f, Args &&...args)
{
  return std::vformat(f, std::make_format_args(args...));
}
}    // namespace fmt
#endif
