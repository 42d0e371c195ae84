#pragma once
#include <memory>
#include <utility>
namespace boost {
using std::shared_ptr;
template <class T, class... A> inline std::shared_ptr<T> make_shared(A &&...a) { return std::make_shared<T>(std::forward<A>(a)...); }
}
