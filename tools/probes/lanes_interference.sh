# timing experiment: which part of a pass bounds the step?  MTL_DBG_SKIP: vgg = no VGG kernels, tf = only VGG kernels,
# w = no parameter-gradient side work, x = bare activation chain (no VGG, no parameter gradients)
for br in 1 0; do for skip in x none; do for l in 1 2 3; do
  r=$(MTL_BRANCHES=$br MTL_DBG_SKIP=$skip python bench.py --no-cpu-baseline --no-gpu-baseline --no-roofline --lanes $l 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f ms/step, %d launches' % (d['ms_per_step'], d['gpu_launches']/d['steps']))")
  echo "branches=$br skip=$skip lanes=$l $r"
done; done; done
