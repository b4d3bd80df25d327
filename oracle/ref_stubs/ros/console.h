// Stand-in for <ros/console.h> (TEST INFRASTRUCTURE ONLY): the reference only logs through these.
#ifndef QPB_ROS_CONSOLE_STANDIN
#define QPB_ROS_CONSOLE_STANDIN
#include <iostream>
#include <sstream>
extern int qpb_ref_error_count;  // bumped by every ROS_ERROR the reference raises (ref_glue.cpp reads it)
#define ROS_ERROR_STREAM_NAMED(name, args) do { qpb_ref_error_count++; } while (0)
#define ROS_WARN_STREAM_NAMED(name, args) do { } while (0)
#define ROS_INFO_STREAM_NAMED(name, args) do { } while (0)
#define ROS_DEBUG_STREAM_NAMED(name, args) do { } while (0)
#define ROS_ERROR_NAMED(name, ...) do { qpb_ref_error_count++; } while (0)
#define ROS_WARN_NAMED(name, ...) do { } while (0)
#define ROS_INFO_NAMED(name, ...) do { } while (0)
#define ROS_DEBUG_NAMED(name, ...) do { } while (0)
#endif
