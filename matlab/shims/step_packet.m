function Pout = step_packet(P, U, GradU, C0, f, dx, dy, dt) %#ok<INUSL>
% Pout = step_packet(P, U, GradU, C0, f, dx, dy, dt)  -- ray_trace_sw/step_packet.m
% P may be a struct ARRAY of packets: all of them advance in one kernel launch.
Pout = swrt_step_packets(P, U, GradU, [], C0, f, dx, dt, 1);
end
