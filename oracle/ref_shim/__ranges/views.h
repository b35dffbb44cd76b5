// common/typedef.h:10 includes this libc++-internal header; libstdc++ has no such file and <ranges> already provides std::views.
#pragma once
#include <ranges>
