// TEST INFRASTRUCTURE ONLY.  The reference's transport logs through the external CUAUV `auvlog`
// library (lib/camera_message_framework.cpp:5); this stand-in swallows the messages so the
// reference's camera message framework can be compiled unmodified by oracle/Makefile.
#pragma once
#define auvlog_info(x) ((void)(x))
#define auvlog_log_stdout(tree, msg) ((void)0)
#define auvlog_log(tree, msg) ((void)0)
