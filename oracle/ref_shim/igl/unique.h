// stand-in: iter_rounding_init.h:9 includes igl/unique.h without using it
#pragma once
