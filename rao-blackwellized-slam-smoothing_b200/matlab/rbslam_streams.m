function opts = rbslam_streams(opts, desc, N_P, N_T, N_K, lastParticleRef)
%RBSLAM_STREAMS  Compat mode: pre-draw rand/randn in the reference's consumption order.
% The reference draws, for every step t>=2 and particle i, one rand (inside sample,
% tools/sample.m:31) followed by the randn's of dynModel (src/particleFilter.m:104-109);
% in smoother sweeps k>=2 the N_P'th particle only draws one rand (src/particleSmoother.m:241),
% and every sweep ends with one rand (:346).  Drawing them here in the same order
% leaves MATLAB's global stream exactly where the reference would leave it.
  switch desc.family
    case 'denseMag3D',     nz = 6;
    case 'denseRadio2D',   nz = 1;
    case 'sparseVisual2D', nz = 3;
  end
  U = zeros(N_P, N_T, N_K); Z = zeros(nz, N_P, N_T, N_K); Uend = zeros(N_K, 1);
  for k = 1:N_K
    for t = 2:N_T
      for i = 1:N_P
        U(i,t,k) = rand;
        if ~(lastParticleRef && k > 1 && i == N_P)
          if strcmp(desc.family, 'denseRadio2D'), Z(:,i,t,k) = randn;
          else, Z(1:3,i,t,k) = randn(3,1); if nz == 6, Z(4:6,i,t,k) = randn(3,1); end
          end
        end
      end
    end
    if lastParticleRef, Uend(k) = rand; end
  end
  opts.U = U; opts.Z = Z; opts.Uend = Uend;
end
