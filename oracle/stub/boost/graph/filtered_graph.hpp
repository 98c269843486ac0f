// see adjacency_list.hpp (one shim for the three Boost.Graph headers the reference includes)
#pragma once
#include "adjacency_list.hpp"
