// stand-in: iter_rounding_init.h:8 includes igl/local_basis.h without using it
#pragma once
