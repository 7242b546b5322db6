# timing experiment: which part of a pass bounds the step?  MTL_DBG_SKIP: vgg = no VGG kernels, tf = only VGG kernels,
# w = no parameter-gradient side work, x = bare activation chain (no VGG, no parameter gradients)
for merge in 0 1; do for skip in none vgg w x; do for l in 1 3; do
  r=$(MTL_MERGE_LOWRANK=$merge MTL_DBG_SKIP=$skip python bench.py --no-cpu-baseline --no-roofline --lanes $l 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f ms/step, %d launches' % (d['ms_per_step'], d['gpu_launches']/d['steps']))")
  echo "merge=$merge skip=$skip lanes=$l $r"
done; done; done
