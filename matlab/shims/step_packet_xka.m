function Pout = step_packet_xka(P, U, GradU, H, C0, f, dx, dy, dt) %#ok<INUSL>
% Pout = step_packet_xka(P, U, GradU, H, C0, f, dx, dy, dt)  -- ray_trace_sw/step_packet_xka.m
Pout = swrt_step_packets(P, U, GradU, H, C0, f, dx, dt, 2);
end
