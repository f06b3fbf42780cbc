#ifndef QDLDL_TYPES_H
#define QDLDL_TYPES_H
#include <limits.h>
typedef int QDLDL_int;
typedef float QDLDL_float;
typedef unsigned char QDLDL_bool;
#define QDLDL_INT_MAX INT_MAX
#define QDLDL_FLOAT
#endif
