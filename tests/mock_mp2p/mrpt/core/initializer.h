#pragma once
// mock: MRPT_INITIALIZER as used at module/src/register.cpp:40
#define MRPT_INITIALIZER(f) static void f(); namespace { struct f##_runner { f##_runner() { f(); } } f##_instance; } static void f()
