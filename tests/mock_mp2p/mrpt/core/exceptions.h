#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#define THROW_EXCEPTION(msg) throw std::runtime_error(msg)
#define THROW_EXCEPTION_FMT(fmt, ...) do { char b__[512]; std::snprintf(b__, sizeof(b__), fmt, __VA_ARGS__); throw std::runtime_error(b__); } while (0)
#define ASSERT_(c) do { if (!(c)) throw std::runtime_error("assert: " #c); } while (0)
