/* TEST INFRASTRUCTURE (oracle).  Link-time stand-ins for the micro-ROS entry points that only
 * create_microros_entities() / destroy_microros_entities() / prepare_task() of RM_task_main.cpp reference; the
 * harness never calls those functions, the symbols merely have to resolve when the library is loaded. */
void rcl_context_get_rmw_context(void) {}
void rcl_node_fini(void) {}
void rcl_publisher_fini(void) {}
void rcl_service_fini(void) {}
void rcl_subscription_fini(void) {}
void rclc_executor_add_service(void) {}
void rclc_executor_add_subscription(void) {}
void rclc_executor_fini(void) {}
void rclc_executor_get_zero_initialized_executor(void) {}
void rclc_executor_init(void) {}
void rclc_node_init_default(void) {}
void rclc_publisher_init_best_effort(void) {}
void rclc_service_init_default(void) {}
void rclc_subscription_init_default(void) {}
void rclc_support_fini(void) {}
void rclc_support_init(void) {}
void rcutils_get_default_allocator(void) {}
void rmw_uros_ping_agent(void) {}
void rmw_uros_set_context_entity_destroy_session_timeout(void) {}
void rmw_uros_sync_session(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__geometry_msgs__msg__Twist(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__ArmInfo(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__CamAngleOrder(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__Command(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__MecanumCommand(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__TimeAngle(void) {}
void rosidl_typesupport_c__get_message_type_support_handle__interfaces__msg__VehicleInfo(void) {}
void rosidl_typesupport_c__get_service_type_support_handle__interfaces__srv__ProcStatus(void) {}
